// C++ host example: the alpha x Cirr sweep of BASELINE config #2, light curves to stdout — through an ensemble handle on one GPU,
// then again in ONE call over every GPU of the box (fb200::sweep: construction streamed behind the kernels) and checked to be
// bit-identical.
//   g++ -O2 -std=c++17 examples/ensemble_sweep.cpp -o build/ensemble_sweep freddi_b200/libfreddi_b200.so -Wl,-rpath,$PWD/freddi_b200
#include <cmath>
#include <cstdio>

#include "../freddi_b200/cpp/freddi_ensemble.hpp"

int main(int argc, char** argv) {
    const int n_alpha = argc > 1 ? std::atoi(argv[1]) : 4, n_cirr = argc > 2 ? std::atoi(argv[2]) : 3;
    const double Msun = 1.98892e33, day = 86400., kpc = 1000. * 3.08567758135e18;
    std::vector<fb200::Arguments> args;
    for (int i = 0; i < n_alpha; ++i)
        for (int j = 0; j < n_cirr; ++j) {
            fb200::Arguments a;
            a.alpha = 0.1 + 0.9 * i / std::max(1, n_alpha - 1);
            a.Cirr = 1e-5 * std::pow(500., double(j) / std::max(1, n_cirr - 1));
            a.Mx = 5 * Msun; a.Mopt = 0.5 * Msun; a.period = 0.25 * day; a.F0 = 2e38; a.distance = 10 * kpc;
            a.time = 50 * day; a.tau = 0.25 * day;
            a.initialcond = FB200_IC_QUASISTAT; a.Thot = 1e4; a.boundcond = FB200_BOUNDCOND_TIRR;
            a.angulardistdisk = FB200_ANGDIST_ISOTROPIC;
            args.push_back(a);
        }
    try {
        fb200::Config cfg;
        cfg.lambda(5e-5);
        fb200::FreddiEnsemble ens(args, cfg);
        ens.evolve();
        const auto rows = ens.rows();
        const auto st = ens.status();
        std::printf("# %zu models, %zu rows x %zu columns each, device time %.2f ms\n", ens.size(), ens.n_rows(), ens.n_cols(),
                    ens.last_evolve_ms());
        for (size_t k = 0; k < ens.size(); ++k) {
            const double* last = &rows[(k * ens.n_rows() + st.rows_valid[k] - 1) * ens.n_cols()];
            std::printf("model %zu alpha=%.3f Cirr=%.2e rows=%d t=%.2f d Mdot=%.4e Lx=%.4e Fnu0=%.4e%s\n", k, args[k].alpha, args[k].Cirr,
                        st.rows_valid[k], last[FB200_COL_T], last[FB200_COL_MDOT], last[FB200_COL_LX], last[FB200_N_BH_COLS],
                        st.status[k] == FB200_MODEL_COLLAPSED ? "  (Rout <= Rin)" : "");
        }
        // the same sweep, one call, all visible GPUs (model k on device k mod G)
        const fb200::SweepResult all = fb200::sweep(args, cfg);
        bool same = all.rows.size() == rows.size();
        for (size_t i = 0; same && i < rows.size(); ++i) same = (rows[i] == all.rows[i]) || (rows[i] != rows[i] && all.rows[i] != all.rows[i]);
        std::printf("# fb200::sweep over %d device(s): %.2f ms for the whole call (%.2f ms on the devices), %lld model-steps, rows %s\n",
                    all.stats.n_devices, all.stats.total_ms, all.stats.device_ms_max, (long long)all.stats.model_steps,
                    same ? "identical" : "DIFFER");
        if (!same) return 2;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
