// TEST INFRASTRUCTURE — not part of the product.  C entry points over the command-line front-end
// (freddi_b200/cpp/freddi_options.hpp, freddi_output.hpp) so that the CPU test-suite can check option parsing and
// the freddi.dat writer byte for byte against the reference's golden files, with rows supplied by the oracle.
#include <cstring>
#include <sstream>

#include "../../freddi_b200/cpp/freddi_options.hpp"
#include "../../freddi_b200/cpp/freddi_output.hpp"

using namespace fb200::cli;

namespace {
int fail(char* err, size_t err_len, const std::exception& e) {
    if (err && err_len) std::snprintf(err, err_len, "%s", e.what());
    return -1;
}
size_t put(const std::string& s, char* out, size_t cap) {
    if (out && cap) { std::memcpy(out, s.data(), std::min(cap, s.size())); }
    return s.size();
}
}  // namespace

extern "C" {

// argv (incl. argv[0]) -> fb200_args_t + ensemble-wide outputs.  lambdas_cm holds up to 16 values.
int cli_parse(int ns, int argc, const char* const* argv, fb200_args_t* args, double* lambdas_cm, int* n_lambdas, int* cold_disk,
              unsigned* precision, unsigned* tempsparsity, int* n_passbands, char* err, size_t err_len) {
    try {
        VariablesMap vm;
        if (!parse_options(ns != 0, argc, argv, vm, nullptr, "/nonexistent")) return 1;
        const RunSetup s = make_setup(vm, ns != 0);
        *args = s.args;
        *n_lambdas = int(s.lambdas.size());
        for (size_t i = 0; i < s.lambdas.size() && i < 16; ++i) lambdas_cm[i] = s.lambdas[i];
        *cold_disk = s.cold_disk;
        *precision = s.precision;
        *tempsparsity = s.tempsparsity;
        *n_passbands = int(s.passband_files.size());
        return 0;
    } catch (const std::exception& e) {
        return fail(err, err_len, e);
    }
}

// The whole freddi.dat for one model: header + rows[n_rows][ncol].  Returns the length (call again with a larger buffer if > cap).
long cli_render_dat(int ns, int argc, const char* const* argv, const char* const* passband_names, int n_passband_names,
                    const double* rows, long n_rows, double rout_cm, double risco_cm, double tau_s, char* out, size_t cap, char* err,
                    size_t err_len) {
    try {
        VariablesMap vm;
        if (!parse_options(ns != 0, argc, argv, vm, nullptr, "/nonexistent")) return -2;
        const RunSetup s = make_setup(vm, ns != 0);
        std::vector<std::string> names(passband_names, passband_names + n_passband_names);
        const std::vector<FieldMeta> fields = short_fields(ns != 0, s.cold_disk, s.lambdas, names, s.star_flux);
        std::ostringstream os;
        const double alphacold = std::isnan(s.args.alphacold) ? s.args.alpha / 10.0 : s.args.alphacold;
        write_dat_header(os, fields, vm, s.precision, alphacold, rout_cm, risco_cm, tau_s);
        for (long r = 0; r < n_rows; ++r)
            if (r % long(s.tempsparsity) == 0) write_row(os, rows + r * long(fields.size()), fields.size(), s.precision);
        return long(put(os.str(), out, cap));
    } catch (const std::exception& e) {
        return fail(err, err_len, e);
    }
}

// One --fulldata file (cpp/src/output.cpp:142-162) from column-major fields[n_fields][Nx]
long cli_render_fulldata(int ns, const double* fields, int n_fields, long Nx, double t_days, long first, long last, unsigned precision,
                         char* out, size_t cap) {
    const std::vector<LongField> lf = disk_structure_fields(ns != 0);
    if (int(lf.size()) != n_fields) return -1;
    std::vector<const double*> cols;
    for (int j = 0; j < n_fields; ++j) cols.push_back(fields + long(j) * Nx);
    std::ostringstream os;
    write_disk_structure(os, lf, cols, t_days, first, last, precision);
    return long(put(os.str(), out, cap));
}

int cli_n_sweep_models(int ns, int argc, const char* const* argv, double* first_values, int cap, char* err, size_t err_len) {
    try {
        VariablesMap vm;
        if (!parse_options(ns != 0, argc, argv, vm, nullptr, "/nonexistent")) return -2;
        const std::vector<Sweep> sw = parse_sweeps(vm, Parser(ns != 0));
        int n = 1;
        for (const Sweep& s : sw) n *= int(s.values.size());
        if (!sw.empty()) for (size_t i = 0; i < sw[0].values.size() && int(i) < cap; ++i) first_values[i] = sw[0].values[i];
        return n;
    } catch (const std::exception& e) {
        return fail(err, err_len, e);
    }
}

}  // extern "C"
