// C++ host of the B200 ensemble engine (header-only, over the C ABI of include/freddi_b200.h).
//
// Mirrors the reference's C++ surface for the hot path:
//   fb200::Arguments            <- FreddiArguments / FreddiNeutronStarArguments (cpp/include/arguments.hpp:336-354,
//                                  cpp/include/ns/ns_arguments.hpp:143-165), CGS units, same member names
//   fb200::FreddiEnsemble       <- N x FreddiEvolution / FreddiNeutronStarEvolution: step(), evolve(), the row of
//                                  FreddiFileOutput::shortDump (cpp/src/output.cpp:197-233), the array getters
//   exceptions                  <- std::invalid_argument / std::runtime_error / std::logic_error at construction
//                                  exactly as the reference's constructors throw; RadiusCollapse is a per-model
//                                  status (the reference's driver catches it per model, application.hpp:31-42)
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/freddi_b200.h"

namespace fb200 {

class RadiusCollapseException : public std::exception {  // cpp/include/exceptions.hpp:6-11
public:
    const char* what() const noexcept override { return "Rout <= Rin"; }
};

inline void check(int rc) {
    if (rc == FB200_OK) return;
    const std::string msg = fb200_last_error();
    switch (rc) {
        case FB200_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
        case FB200_ERR_LOGIC: throw std::logic_error(msg);
        default: throw std::runtime_error(msg);
    }
}

struct Arguments : fb200_args_t {
    explicit Arguments(bool neutron_star = false) { fb200_args_default(this, neutron_star ? 1 : 0); }
};

struct Config : fb200_config_t {
    Config() { fb200_config_default(this); }
    Config& lambda(double cm) {
        if (n_lambdas >= FB200_MAX_LAMBDAS) throw std::invalid_argument("too many lambdas");
        lambdas[n_lambdas++] = cm;
        return *this;
    }
};

class FreddiEnsemble {
public:
    FreddiEnsemble(const std::vector<Arguments>& args, const Config& cfg = Config()) {
        std::vector<fb200_args_t> raw(args.begin(), args.end());
        size_t bad = 0;
        check(fb200_ensemble_create(raw.data(), raw.size(), &cfg, &h_, &bad));
    }
    ~FreddiEnsemble() { fb200_ensemble_destroy(h_); }
    FreddiEnsemble(const FreddiEnsemble&) = delete;
    FreddiEnsemble& operator=(const FreddiEnsemble&) = delete;

    size_t size() const { return fb200_ensemble_n_models(h_); }
    size_t Nx() const { return fb200_ensemble_nx(h_); }
    size_t n_rows() const { return fb200_ensemble_n_rows(h_); }
    size_t n_cols() const { return fb200_ensemble_n_cols(h_); }

    // FreddiEvolution::step() for every model that has steps left (cpp/include/freddi_evolution.hpp:52-53)
    void step() { check(fb200_ensemble_evolve(h_, 1)); }
    // the whole driver loop of cpp/include/application.hpp:25-43 in one launch
    void evolve(int64_t n_steps = -1) { check(fb200_ensemble_evolve(h_, n_steps)); }
    float last_evolve_ms() { float ms = 0; check(fb200_ensemble_last_evolve_ms(h_, &ms)); return ms; }

    // rows[model][row][col], NaN where a model never got (collapsed earlier / shorter run)
    std::vector<double> rows() {
        std::vector<double> out(size() * n_rows() * n_cols());
        check(fb200_ensemble_rows(h_, out.data(), out.size()));
        return out;
    }
    struct Status { std::vector<int32_t> status, rows_valid; std::vector<int64_t> picard; };
    Status status() {
        Status s{std::vector<int32_t>(size()), std::vector<int32_t>(size()), std::vector<int64_t>(size())};
        check(fb200_ensemble_status(h_, s.status.data(), s.rows_valid.data(), s.picard.data()));
        return s;
    }
    // throws like FreddiEvolution::step() would have for model k
    void throw_if_collapsed(size_t k) { if (status().status.at(k) == FB200_MODEL_COLLAPSED) throw RadiusCollapseException(); }
    std::vector<double> field(int which, int64_t row = -1) {  // FB200_FIELD_*: h(), R(), F(), Sigma(), Tph(), ...
        std::vector<double> out(size() * Nx());
        check(fb200_ensemble_field(h_, which, row, out.data(), out.size()));
        return out;
    }
    std::vector<double> flux(const std::vector<double>& lambdas, bool cold = false, int64_t row = -1) {
        std::vector<double> out(size() * lambdas.size());
        check(fb200_ensemble_flux(h_, cold ? 1 : 0, row, lambdas.data(), lambdas.size(), out.data(), out.size()));
        return out;
    }
    fb200_ensemble_t* handle() { return h_; }

private:
    fb200_ensemble_t* h_ = nullptr;
};

// The same over several GPUs of this process (fb200_sharded_*): model k lives on device k mod G, all devices evolve at once,
// rows() gathers into one array.  devices empty = all visible.
class ShardedEnsemble {
public:
    ShardedEnsemble(const std::vector<Arguments>& args, const Config& cfg = Config(), const std::vector<int32_t>& devices = {}) {
        std::vector<fb200_args_t> raw(args.begin(), args.end());
        size_t bad = 0;
        check(fb200_sharded_create(raw.data(), raw.size(), &cfg, devices.empty() ? nullptr : devices.data(), int32_t(devices.size()), &h_, &bad));
    }
    ~ShardedEnsemble() { fb200_sharded_destroy(h_); }
    ShardedEnsemble(const ShardedEnsemble&) = delete;
    ShardedEnsemble& operator=(const ShardedEnsemble&) = delete;

    size_t size() const { return fb200_sharded_n_models(h_); }
    size_t Nx() const { return fb200_sharded_nx(h_); }
    size_t n_rows() const { return fb200_sharded_n_rows(h_); }
    size_t n_cols() const { return fb200_sharded_n_cols(h_); }
    int n_shards() const { return fb200_sharded_n_shards(h_); }
    void step() { check(fb200_sharded_evolve(h_, 1)); }
    void evolve(int64_t n_steps = -1) { check(fb200_sharded_evolve(h_, n_steps)); }
    void reset() { check(fb200_sharded_reset(h_)); }
    float last_evolve_ms() { float ms = 0; check(fb200_sharded_last_evolve_ms(h_, &ms)); return ms; }
    std::vector<double> rows() {
        std::vector<double> out(size() * n_rows() * n_cols());
        check(fb200_sharded_rows(h_, out.data(), out.size()));
        return out;
    }
    FreddiEnsemble::Status status() {
        FreddiEnsemble::Status s{std::vector<int32_t>(size()), std::vector<int32_t>(size()), std::vector<int64_t>(size())};
        check(fb200_sharded_status(h_, s.status.data(), s.rows_valid.data(), s.picard.data()));
        return s;
    }
    std::vector<double> field(int which, int64_t row = -1) {
        std::vector<double> out(size() * Nx());
        check(fb200_sharded_field(h_, which, row, out.data(), out.size()));
        return out;
    }
    fb200_sharded_t* handle() { return h_; }
    fb200_ensemble_t* shard(int k) { return fb200_sharded_shard(h_, k); }

private:
    fb200_sharded_t* h_ = nullptr;
};

// The whole driver of the reference (cpp/include/application.hpp:17-43) for every model in one call: argument structs in,
// light curves out, construction streamed behind the evolve kernels of all `devices` (fb200_sweep_run).
struct SweepResult {
    std::vector<double> rows;  // [model][row][col]
    std::vector<int32_t> status, rows_valid;
    std::vector<int64_t> picard;
    std::vector<double> t_end;  // FreddiState::t() where each model stopped [s]
    fb200_sweep_stats_t stats{};
    size_t n_rows = 0, n_cols = 0;
};
inline SweepResult sweep(const std::vector<Arguments>& args, const Config& cfg = Config(), const std::vector<int32_t>& devices = {}) {
    std::vector<fb200_args_t> raw(args.begin(), args.end());
    SweepResult r;
    check(fb200_sweep_plan(raw.data(), raw.size(), &cfg, &r.n_rows, &r.n_cols));
    r.rows.resize(raw.size() * r.n_rows * r.n_cols);
    r.status.resize(raw.size());
    r.rows_valid.resize(raw.size());
    r.picard.resize(raw.size());
    r.t_end.resize(raw.size());
    size_t bad = 0;
    check(fb200_sweep_run(raw.data(), raw.size(), &cfg, devices.empty() ? nullptr : devices.data(), int32_t(devices.size()), r.rows.data(),
                          r.rows.size(), r.status.data(), r.rows_valid.data(), r.picard.data(), r.t_end.data(), &r.stats, &bad));
    return r;
}

}  // namespace fb200
