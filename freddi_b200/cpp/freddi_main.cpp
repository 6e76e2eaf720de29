// `freddi-b200` / `freddi-ns-b200`: the reference's two executables (cpp/main.cpp, cpp/main-ns.cpp,
// run_application in cpp/include/application.hpp:16-45) over the B200 ensemble engine.  Same options, same INI search
// path, same freddi.dat / --fulldata files; the evolution itself is one launch of the sm_100a kernel through the C ABI
// (there is no CPU path: without a CUDA device the run fails).
//
//   freddi-b200 --alpha=0.25 --Mx=5 ... [--sweep alpha=0.1:1.0:40 --sweep Cirr=1e-5:5e-3:25:log]
//
// Without --sweep one model is evolved and PREFIX.dat is what `freddi` writes.  With sweeps the cartesian product is
// evolved at once and model k goes to PREFIX-k.dat, each a complete freddi.dat of that model (its own values of the
// swept options in the "### Parameters" echo).
//
// The neutron-star variant is selected by the executable name (contains "-ns") or by -DFREDDI_NS.
#include <cmath>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>

#include "freddi_ensemble.hpp"
#include "freddi_options.hpp"
#include "freddi_output.hpp"

namespace {

using namespace fb200;
using namespace fb200::cli;

struct PassbandData {
    std::string name;
    std::vector<double> lambdas, transmissions;
};

// EnergyPassband(filepath, PhotonCounter): cpp/src/passband.cpp:10-43, cpp/src/options.cpp:308-322
PassbandData load_passband(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw OptionError("EnergyPassband file doesn't exist");
    PassbandData pb;
    for (double lam, tr; f >> lam >> tr;) {
        lam = angstromToCm(lam);
        pb.lambdas.push_back(lam);
        pb.transmissions.push_back(tr * lam);  // photon counter
    }
    std::string name = path;  // nameFromPath: for (; !path.extension().empty(); path = path.stem());
    for (;;) {
        const size_t slash = name.find_last_of('/');
        const std::string base = slash == std::string::npos ? name : name.substr(slash + 1);
        const size_t dot = base.find_last_of('.');
        if (dot == std::string::npos || dot == 0) break;
        name = base.substr(0, dot);
    }
    pb.name = name;
    return pb;
}

}  // namespace

int main(int ac, char** av) {
#ifdef FREDDI_NS
    const bool ns = true;
#else
    const bool ns = std::strstr(av[0] ? av[0] : "", "-ns") != nullptr;
#endif
    VariablesMap vm;
    std::vector<Sweep> sweeps;
    try {
        std::string help;
        if (!parse_options(ns, ac, av, vm, &help)) {
            std::cout << help << std::endl;
            return 1;  // run_application returns false after --help (cpp/include/options.hpp:112-115, cpp/main.cpp:8-10)
        }
        sweeps = parse_sweeps(vm, Parser(ns));
    } catch (const OptionError& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }

    // ---- the models: cartesian product of the sweeps (first sweep = slowest axis) ----
    size_t n_models = 1;
    for (const Sweep& s : sweeps) n_models *= s.values.size();
    std::vector<VariablesMap> vms;
    std::vector<RunSetup> setups;
    try {
        for (size_t k = 0; k < n_models; ++k) {
            VariablesMap m = vm;
            size_t rem = k;
            for (size_t j = sweeps.size(); j-- > 0;) {
                const Sweep& s = sweeps[j];
                Value v;
                v.kind = Kind::Double;
                v.d = s.values[rem % s.values.size()];
                rem /= s.values.size();
                m[s.name] = v;
            }
            if (!sweeps.empty()) m["prefix"].s = vm.at("prefix").s + "-" + std::to_string(k);
            setups.push_back(make_setup(m, ns));
            vms.push_back(std::move(m));
        }
    } catch (const OptionError& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    }
    const RunSetup& s0 = setups.front();

    try {
        std::vector<PassbandData> passbands;
        for (const std::string& path : s0.passband_files) passbands.push_back(load_passband(path));
        Config cfg;
        for (double l : s0.lambdas) cfg.lambda(l);
        cfg.cold_disk = s0.cold_disk ? 1 : 0;
        cfg.record_F = s0.fulldata ? 1 : 0;
        cfg.star_flux = s0.star_flux ? 1 : 0;
        if (passbands.size() > FB200_MAX_PASSBANDS) throw OptionError("at most " + std::to_string(FB200_MAX_PASSBANDS) + " --passband files");
        cfg.n_passbands = int(passbands.size());
        std::vector<std::string> passband_names;
        for (size_t p = 0; p < passbands.size(); ++p) {
            cfg.passband_len[p] = int(passbands[p].lambdas.size());
            cfg.passband_lambdas[p] = passbands[p].lambdas.data();
            cfg.passband_transmissions[p] = passbands[p].transmissions.data();
            passband_names.push_back(passbands[p].name);
        }
        // The reference's driver loops i_t = 0 .. int(time / tau) INCLUSIVE and steps after every dump (application.hpp:25-43): one
        // step more than it dumps, whose only visible effect is the "terminated prematurely" line when that step collapses.
        // The models are therefore evolved for time + tau (tau made explicit: its default is time / 200); rows are dumped for the
        // requested time only.
        std::vector<Arguments> args(n_models, Arguments(ns));
        std::vector<long> n_dump(n_models);
        for (size_t k = 0; k < n_models; ++k) {
            fb200_model_info_t pre;
            check(fb200_args_resolve(&setups[k].args, &pre, nullptr, nullptr));
            n_dump[k] = long(static_cast<int>(setups[k].args.time / pre.tau)) + 1;  // i_t = 0 .. int(time / tau)
            static_cast<fb200_args_t&>(args[k]) = setups[k].args;
            args[k].tau = pre.tau;
            args[k].time = setups[k].args.time + pre.tau;
        }

        // --devices all | 0,1,...: the models are spread over several GPUs of the box from this one process (model k -> device
        // k mod G).  Light curves only: the streamed one-call sweep; with --fulldata the sharded handle (per-row structure read-out).
        std::vector<int32_t> devices;
        const bool multi = vm.count("devices") != 0;
        if (multi && vm.at("devices").s != "all") {
            std::string tok;
            for (char c : vm.at("devices").s + ",") {
                if (c == ',') {
                    if (!tok.empty()) devices.push_back(int32_t(Parser::to_uint(tok, "devices")));
                    tok.clear();
                } else {
                    tok += c;
                }
            }
            if (devices.empty()) throw OptionError("--devices expects 'all' or a comma-separated list of CUDA ordinals");
        }
        std::unique_ptr<FreddiEnsemble> ens;
        std::unique_ptr<ShardedEnsemble> sharded;
        std::vector<fb200_model_info_t> info(n_models);
        std::vector<double> rows;
        FreddiEnsemble::Status st;
        std::vector<double> t_end(n_models);  // FreddiState::t() where each model stopped [s]
        size_t ncol = 0, nrows = 0, Nx = size_t(setups[0].args.Nx);
        if (multi && !s0.fulldata) {
            SweepResult r = sweep(args, cfg, devices);
            rows.swap(r.rows);
            t_end.swap(r.t_end);
            st = FreddiEnsemble::Status{std::move(r.status), std::move(r.rows_valid), std::move(r.picard)};
            ncol = r.n_cols;
            nrows = r.n_rows;
            for (size_t k = 0; k < n_models; ++k) check(fb200_args_resolve(&args[k], &info[k], nullptr, nullptr));
        } else if (multi) {
            sharded.reset(new ShardedEnsemble(args, cfg, devices));
            check(fb200_sharded_info(sharded->handle(), info.data()));
            sharded->evolve(-1);
            rows = sharded->rows();
            st = sharded->status();
            check(fb200_sharded_state(sharded->handle(), t_end.data(), nullptr, nullptr, nullptr));
            ncol = sharded->n_cols(); nrows = sharded->n_rows(); Nx = sharded->Nx();
        } else {
            ens.reset(new FreddiEnsemble(args, cfg));
            check(fb200_ensemble_info(ens->handle(), info.data()));
            ens->evolve(-1);
            rows = ens->rows();
            st = ens->status();
            check(fb200_ensemble_state(ens->handle(), t_end.data(), nullptr, nullptr, nullptr));
            ncol = ens->n_cols(); nrows = ens->n_rows(); Nx = ens->Nx();
        }

        const std::vector<FieldMeta> fields = short_fields(ns, s0.cold_disk, s0.lambdas, passband_names, s0.star_flux);
        if (fields.size() != ncol) throw std::logic_error("column count of the engine and of the writer differ");
        const std::vector<LongField> long_fields = disk_structure_fields(ns);

        long max_dump = 0;
        for (size_t k = 0; k < n_models; ++k) {
            if (n_dump[k] > long(nrows)) n_dump[k] = long(nrows);
            max_dump = std::max(max_dump, n_dump[k]);
        }

        std::vector<std::unique_ptr<std::ofstream>> files(n_models);
        for (size_t k = 0; k < n_models; ++k) {
            const RunSetup& s = setups[k];
            std::ostream* os = &std::cout;
            if (!s.to_stdout) {
                files[k].reset(new std::ofstream(s.dir + "/" + s.prefix + ".dat"));
                os = files[k].get();
            }
            const double alphacold = std::isnan(s.args.alphacold) ? s.args.alpha / 10.0 : s.args.alphacold;  // arguments.hpp:84
            write_dat_header(*os, fields, vms[k], s.precision, alphacold, info[k].rout, info[k].risco, info[k].tau);
            for (long r = 0; r < n_dump[k] && r < st.rows_valid[k]; ++r)
                if (r % long(s.tempsparsity) == 0) write_row(*os, &rows[(k * nrows + size_t(r)) * ncol], ncol, s.precision);
            os->flush();
            if (st.status[k] == FB200_MODEL_COLLAPSED && st.rows_valid[k] <= n_dump[k]) {
                // The step from state i_t = rows_valid - 1 threw RadiusCollapseException.  freddi->t() already holds the tau
                // FreddiState::step added when the outer radius collapsed in truncateOuterRadius (freddi_state.cpp:147-153), not
                // when the NS inner radius did (truncateInnerRadius runs first); std::cerr prints it at its default precision.
                std::cerr << "Freddi terminated prematurely, i_t = " << st.rows_valid[k] - 1 << ", t = " << t_end[k] / 86400.
                          << " (days), reason: Rout <= Rin";
                if (n_models > 1) std::cerr << " [model " << k << "]";
                std::cerr << std::endl;
            }
        }

        if (s0.fulldata) {  // PREFIX_<i_t>.dat, cpp/src/output.cpp:142-162
            std::vector<std::vector<double>> cols(long_fields.size());
            std::vector<int64_t> first(n_models), last(n_models);
            for (long r = 0; r < max_dump; ++r) {
                bool any = false;
                for (size_t k = 0; k < n_models; ++k) any |= (r < n_dump[k] && r < st.rows_valid[k] && r % long(setups[k].tempsparsity) == 0);
                if (!any) continue;
                if (sharded) {
                    for (size_t j = 0; j < long_fields.size(); ++j) cols[j] = sharded->field(long_fields[j].field, r);
                    const size_t G = size_t(sharded->n_shards());
                    for (size_t d = 0; d < G; ++d) {  // row bounds are per device: shard d holds the models d, d + G, ...
                        const size_t nd = (n_models - d + G - 1) / G;
                        std::vector<int64_t> f(nd), l(nd);
                        check(fb200_ensemble_row_bounds(sharded->shard(int(d)), r, f.data(), l.data()));
                        for (size_t j = 0; j < nd; ++j) { first[d + j * G] = f[j]; last[d + j * G] = l[j]; }
                    }
                } else {
                    for (size_t j = 0; j < long_fields.size(); ++j) cols[j] = ens->field(long_fields[j].field, r);
                    check(fb200_ensemble_row_bounds(ens->handle(), r, first.data(), last.data()));
                }
                for (size_t k = 0; k < n_models; ++k) {
                    const RunSetup& s = setups[k];
                    if (!(r < n_dump[k] && r < st.rows_valid[k] && r % long(s.tempsparsity) == 0)) continue;
                    std::ofstream full(s.dir + "/" + s.prefix + "_" + std::to_string(r) + ".dat");
                    std::vector<const double*> colp;
                    for (size_t j = 0; j < long_fields.size(); ++j) colp.push_back(&cols[j][k * Nx]);
                    const long upto = s.cold_disk ? long(Nx) - 1 : long(last[k]);
                    write_disk_structure(full, long_fields, colp, rows[(k * nrows + size_t(r)) * ncol], long(first[k]), upto, s.precision);
                }
            }
        }
    } catch (const OptionError& e) {
        std::cerr << "Error: " << e.what() << std::endl;
        return 1;
    } catch (const std::exception& e) {
        // the reference lets constructor exceptions (std::invalid_argument, std::runtime_error, ...) terminate the process
        std::cerr << "terminate called after throwing an instance of a standard exception\n  what():  " << e.what() << std::endl;
        return 134;
    }
    return 0;
}
