// TEST INFRASTRUCTURE — not part of the product.  Only tests/, __graft_entry__.smoke() and bench.py's
// CPU-baseline legs may load the library built from this file.
//
// C-ABI driver over the UNMODIFIED reference translation units (MIN_SRC + NS_MIN_SRC of the
// reference's CMakeLists.txt:15-45,54-59), compiled in place from /root/reference by oracle/Makefile
// against oracle/boost_shim (Boost is not installed here).  It builds Arguments objects directly, the
// way cpp/pywrap/pywrap_arguments.cpp:15-164 does (no program_options), and runs the reference's own
// driver loop `dump(); step();` (cpp/include/application.hpp:25-43) with the getters behind
// cpp/src/output.cpp:197-292 and cpp/src/ns/ns_output.cpp:6-19 evaluated into double buffers.
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <optional>
#include <string>
#include <thread>
#include <vector>

#include <algorithm>

#include <arguments.hpp>
#include <exceptions.hpp>
#include <freddi_evolution.hpp>
#include <ns/ns_arguments.hpp>
#include <ns/ns_evolution.hpp>
#include <spectrum.hpp>
#include <unit_transformation.hpp>
#include <util.hpp>

#include "../include/freddi_b200.h"

namespace {

std::optional<double> opt(double v) { return std::isnan(v) ? std::optional<double>() : std::optional<double>(v); }

const char* pick(const char* const* table, int n, int v) { return (v >= 0 && v < n) ? table[v] : "<invalid>"; }

const char* const kOpacity[] = {"Kramers", "OPAL"};
const char* const kBound[] = {"Teff", "Tirr"};
const char* const kGrid[] = {"log", "linear"};
const char* const kAng[] = {"plane", "isotropic"};
const char* const kIC[] = {"powerF", "linearF", "powerSigma", "sineF", "gaussF", "quasistat", "quasistat-ns"};
const char* const kWind[] = {"no", "SS73C", "Cambier2013", "__testA__", "__testB__", "__testC__", "ShieldsOscil1986",
                             "Janiuk2015", "Shields1986", "Woods1996AGN", "Woods1996", "toy"};
const char* const kNsprop[] = {"dummy", "sibgatullinsunyaev2000", "newt"};
const char* const kFp[] = {"no-outflow", "propeller", "corotation-block", "eksi-kutlu2010", "romanova2018", "geometrical"};
const char* const kKappat[] = {"const", "corstep", "romanova2018"};
const char* const kRedshift[] = {"off", "on"};

pard wind_params(const fb200_args_t& a) {
    const double* p = a.windparams;
    switch (a.windtype) {
        case FB200_WIND_CAMBIER2013: return {{"kC", p[0]}, {"RIC", p[1]}};
        case FB200_WIND_TESTA: return {{"kA", p[0]}};
        case FB200_WIND_TESTB: return {{"kB", p[0]}};
        case FB200_WIND_TESTC: return {{"kC", p[0]}};
        case FB200_WIND_SHIELDSOSCIL1986: return {{"C_w", p[0]}, {"R_w", p[1]}};
        case FB200_WIND_JANIUK2015: return {{"A_0", p[0]}, {"B_1", p[1]}};
        case FB200_WIND_SHIELDS1986: return {{"Xi_max", p[0]}, {"T_ic", p[1]}, {"Pow", p[2]}};
        case FB200_WIND_WOODS1996AGN: return {{"C_0", p[0]}, {"T_ic", p[1]}};
        case FB200_WIND_WOODS1996: return {{"Xi_max", p[0]}, {"T_ic", p[1]}, {"Pow", p[2]}};
        case FB200_WIND_TOY: return {{"C_w", p[0]}};
        default: return {};
    }
}

// Counts calls of the opacity closure made by the solver (range first+1..last): 1 + P per step.
template <class Base>
class Counting : public Base {
public:
    mutable long solver_wunc_calls = 0;
    template <class A> explicit Counting(const A& a) : Base(a) {}
protected:
    vecd wunction(const vecd& h, const vecd& F, size_t f, size_t l) const override {
        if (f == this->first() + 1) ++solver_wunc_calls;
        return Base::wunction(h, F, f, l);
    }
};

struct Model {
    bool is_ns = false;
    std::shared_ptr<Counting<FreddiEvolution>> bh;
    std::shared_ptr<Counting<FreddiNeutronStarEvolution>> ns;
    std::vector<double> lambdas;
    size_t n_passbands = 0;
    int star_flux = 0;
    long last_picard = 0;
    FreddiEvolution* ev() const { return is_ns ? static_cast<FreddiEvolution*>(ns.get()) : static_cast<FreddiEvolution*>(bh.get()); }
};

// Passbands of the models built next (FluxArguments::passbands, cpp/include/arguments.hpp:287): set by
// ref_set_passbands, read-only while models are built and run.
std::vector<EnergyPassband> g_passbands;
double g_define_mdot0 = NAN;  // ref_define_mdot0: value of the member the reference leaves uninitialised (NaN: leave it)
int g_star_flux = 0;  // ref_set_star_flux: add the Fnu*_star, _star_min, _star_max columns (--starflux, cpp/src/output.cpp:234-255,272-291)

FluxArguments* make_flux(const fb200_args_t& a, const std::vector<double>& lambdas) {
    return new FluxArguments(a.colourfactor, a.emin, a.emax, a.staralbedo, a.inclination, a.ephemerist0, a.distance,
                             false, false, lambdas, g_passbands);
}

Model* build(const fb200_args_t& a, const double* lambdas, int n_lambdas) {
    auto m = std::make_unique<Model>();
    m->is_ns = a.is_ns != 0;
    m->lambdas.assign(lambdas, lambdas + n_lambdas);
    m->n_passbands = g_passbands.size();
    m->star_flux = g_star_flux;
    const std::string opacity = pick(kOpacity, 2, a.opacity), bound = pick(kBound, 2, a.boundcond),
                      ic = pick(kIC, 7, a.initialcond), wind = pick(kWind, 12, a.windtype);
    const double Tirr2Tvishot = std::pow(a.Qirr2Qvishot, 0.25);  // cpp/pywrap/pywrap_freddi_evolution.cpp:148
    auto* general = new GeneralArguments("", "", 0, 1, false, false);
    auto* calc = new CalculationArguments(a.inittime, a.time, opt(a.tau), a.Nx, pick(kGrid, 2, a.gridscale),
                                          static_cast<unsigned short>(a.starlod), a.eps);
    if (!m->is_ns) {
        auto* basic = new BasicDiskBinaryArguments(a.alpha, opt(a.alphacold), a.Mx, a.kerr, a.period, a.Mopt,
                                                   a.rochelobefill, a.Topt, opt(a.rin), opt(a.rout), opt(a.risco));
        auto* disk = new DiskStructureArguments(*basic, opacity, a.Mdotout, bound, a.Thot, Tirr2Tvishot, ic, opt(a.F0),
                                                opt(a.Mdisk0), opt(a.Mdot0), opt(a.powerorder), opt(a.gaussmu),
                                                opt(a.gausssigma), wind, wind_params(a));
        // DiskStructureArguments::Mdot0 is read by Cambier2013Wind (cpp/src/freddi_state.cpp:408,411) but no constructor ever
        // sets it (cpp/src/arguments.cpp:42-55): ref_define_mdot0 gives the member a defined value for the models built next
        if (g_define_mdot0 == g_define_mdot0) disk->Mdot0 = g_define_mdot0;
        auto* irr = new SelfIrradiationArguments(a.Cirr, a.irrindex, a.Cirrcold, a.irrindexcold, a.h2rcold,
                                                 pick(kAng, 2, a.angulardistdisk));
        FreddiArguments args(general, basic, disk, irr, make_flux(a, m->lambdas), calc);
        m->bh = std::make_shared<Counting<FreddiEvolution>>(args);
    } else {
        pard fpparams, kparams;
        if (a.fptype == FB200_FP_ROMANOVA2018) fpparams = {{"par1", a.fpparams[0]}, {"par2", a.fpparams[1]}};
        if (a.fptype == FB200_FP_GEOMETRICAL) fpparams = {{"chi", a.fpparams[0]}};
        if (a.kappattype == FB200_KAPPAT_CONST) kparams = {{"value", a.kappatparams[0]}};
        if (a.kappattype == FB200_KAPPAT_CORSTEP) kparams = {{"in", a.kappatparams[0]}, {"out", a.kappatparams[1]}};
        if (a.kappattype == FB200_KAPPAT_ROMANOVA2018)
            kparams = {{"in", a.kappatparams[0]}, {"out", a.kappatparams[1]}, {"par1", a.kappatparams[2]}, {"par2", a.kappatparams[3]}};
        auto* nsa = new NeutronStarArguments(pick(kNsprop, 3, a.nsprop), opt(a.freqx), opt(a.Rx), a.Bx, a.hotspotarea,
                                             a.epsilonAlfven, a.inversebeta, a.Rdead, pick(kFp, 6, a.fptype), fpparams,
                                             pick(kKappat, 3, a.kappattype), kparams, pick(kRedshift, 2, a.nsgravredshift));
        auto* basic = new NeutronStarBasicDiskBinaryArguments(*nsa, a.alpha, opt(a.alphacold), a.Mx, a.kerr, a.period,
                                                              a.Mopt, a.rochelobefill, a.Topt, opt(a.rin), opt(a.rout),
                                                              opt(a.risco));
        // pywrap builds a plain DiskStructureArguments for NS too (pywrap_freddi_evolution.cpp:245-254); the
        // quasistat-ns initial condition only exists in NeutronStarDiskStructureArguments (ns_options.cpp path).
        DiskStructureArguments* disk;
        if (a.initialcond == FB200_IC_QUASISTAT_NS) {
            disk = new NeutronStarDiskStructureArguments(*nsa, *basic, opacity, a.Mdotout, bound, a.Thot, Tirr2Tvishot, ic,
                                                         opt(a.F0), opt(a.Mdisk0), opt(a.Mdot0), opt(a.powerorder),
                                                         opt(a.gaussmu), opt(a.gausssigma), wind, wind_params(a));
        } else {
            disk = new DiskStructureArguments(*basic, opacity, a.Mdotout, bound, a.Thot, Tirr2Tvishot, ic, opt(a.F0),
                                              opt(a.Mdisk0), opt(a.Mdot0), opt(a.powerorder), opt(a.gaussmu),
                                              opt(a.gausssigma), wind, wind_params(a));
        }
        auto* irr = new NeutronStarSelfIrradiationArguments(a.Cirr, a.irrindex, a.Cirrcold, a.irrindexcold, a.h2rcold,
                                                            pick(kAng, 2, a.angulardistdisk), pick(kAng, 2, a.angulardistns));
        FreddiNeutronStarArguments args(general, basic, disk, irr, make_flux(a, m->lambdas), calc, nsa);
        m->ns = std::make_shared<Counting<FreddiNeutronStarEvolution>>(args);
    }
    return m.release();
}

int n_cols(const Model& m, int cold_disk) {
    return FB200_N_BH_COLS + static_cast<int>(m.lambdas.size() + m.n_passbands) * ((cold_disk ? 2 : 1) + (m.star_flux ? 3 : 0)) +
           (m.is_ns ? FB200_N_NS_COLS : 0);
}

// cpp/src/output.cpp:197-214,219-233 and cpp/src/ns/ns_output.cpp:6-19, same evaluation order.
void fill_row(Model& m, int cold_disk, double* out) {
    FreddiEvolution* f = m.ev();
    int k = 0;
    out[k++] = sToDay(f->t());
    out[k++] = f->Mdot_in();
    out[k++] = f->Mdisk();
    out[k++] = cmToSun(f->R()[f->last()]);
    out[k++] = f->Sigma()[f->last()];
    out[k++] = f->Kirr()[f->last()];
    out[k++] = f->Height()[f->last()] / f->R()[f->last()];
    out[k++] = f->Tph()[f->last()];
    out[k++] = f->Tirr()[f->last()];
    out[k++] = m::pow<4>(f->Tirr()[f->last()] / f->Tph_vis()[f->last()]);
    out[k++] = kToKev(*std::max_element(f->Tph_X().begin() + f->first(), f->Tph_X().begin() + f->last() + 1));
    out[k++] = f->Lx();
    out[k++] = f->Lbol_disk();
    out[k++] = f->Fx();
    out[k++] = f->Lbol_disk() * f->angular_dist_disk(f->cosi()) / (FOUR_M_PI * m::pow<2>(f->distance()));
    for (double lambda : m.lambdas) {
        out[k++] = f->flux(lambda);
        if (cold_disk) out[k++] = f->flux_region<FreddiState::ColdRegion>(lambda);
        if (m.star_flux) {
            out[k++] = f->flux_star(lambda);
            out[k++] = f->flux_star(lambda, 0.0);
            out[k++] = f->flux_star(lambda, M_PI);
        }
    }
    for (const auto& pb : f->args().flux->passbands) {  // cpp/src/output.cpp:258-291
        out[k++] = f->flux(pb);
        if (cold_disk) out[k++] = f->flux_region<FreddiState::ColdRegion>(pb);
        if (m.star_flux) {
            out[k++] = f->flux_star(pb);
            out[k++] = f->flux_star(pb, 0.0);
            out[k++] = f->flux_star(pb, M_PI);
        }
    }
    if (m.is_ns) {
        auto* n = m.ns.get();
        out[k++] = n->R()[n->first()];
        out[k++] = n->eta_ns();
        out[k++] = n->Lx_ns();
        out[k++] = n->Lbol_ns();
        out[k++] = n->Fx_ns();
        out[k++] = n->Lbol_ns() * n->angular_dist_ns(n->cosi()) / (FOUR_M_PI * m::pow<2>(n->distance()));
        out[k++] = kToKev(n->T_hot_spot());
        out[k++] = n->fp();
        out[k++] = n->Fmagn()[n->first()];
        out[k++] = n->F()[n->first()];
    }
}

int classify(const std::exception& e) {
    if (dynamic_cast<const std::invalid_argument*>(&e)) return FB200_ERR_INVALID_ARGUMENT;
    if (dynamic_cast<const std::logic_error*>(&e)) return FB200_ERR_LOGIC;  // invalid_argument is-a logic_error, tested first
    if (dynamic_cast<const std::out_of_range*>(&e)) return FB200_ERR_LOGIC;
    return FB200_ERR_RUNTIME;
}

int do_step(Model& m) {
    auto* bh = m.bh.get();
    auto* ns = m.ns.get();
    long& counter = m.is_ns ? ns->solver_wunc_calls : bh->solver_wunc_calls;
    const long before = counter;
    try {
        m.ev()->step();
    } catch (const RadiusCollapseException&) {
        m.last_picard = counter - before > 0 ? counter - before - 1 : 0;
        return 1;
    }
    m.last_picard = counter - before - 1;
    return 0;
}

}  // namespace

extern "C" {

// Tabulated passbands for the models created afterwards: n tables, lens[p] samples each, wavelengths [cm] and
// transmissions back to back.  n = 0 clears.  t_dnu[p] (optional) returns EnergyPassband::t_dnu.
void ref_set_passbands(int n, const int* lens, const double* lambdas, const double* transmissions, double* t_dnu) {
    g_passbands.clear();
    size_t off = 0;
    for (int p = 0; p < n; ++p) {
        std::vector<EnergyPassband::PassbandPoint> data;
        for (int j = 0; j < lens[p]; ++j) data.emplace_back(lambdas[off + j], transmissions[off + j]);
        off += lens[p];
        g_passbands.emplace_back("pb" + std::to_string(p), data);
        if (t_dnu) t_dnu[p] = g_passbands.back().t_dnu;
    }
}

void ref_set_star_flux(int on) { g_star_flux = on; }

// Companion-star triangles of a model: centre (3), normal (3), area per triangle, [n][7]; returns n (out may be NULL).
long ref_star_triangles(void* h, double* out) {
    const auto& tr = static_cast<Model*>(h)->ev()->star().triangles();
    if (out)
        for (size_t i = 0; i < tr.size(); ++i) {
            double* o = out + 7 * i;
            o[0] = tr[i].center().x(); o[1] = tr[i].center().y(); o[2] = tr[i].center().z();
            o[3] = tr[i].normal().x(); o[4] = tr[i].normal().y(); o[5] = tr[i].normal().z();
            o[6] = tr[i].area();
        }
    return static_cast<long>(tr.size());
}

void* ref_create(const fb200_args_t* a, const double* lambdas, int n_lambdas, int* err_code, char* err, size_t err_len) {
    try {
        if (err_code) *err_code = 0;
        return build(*a, lambdas, n_lambdas);
    } catch (const std::exception& e) {
        if (err_code) *err_code = classify(e);
        if (err && err_len) std::snprintf(err, err_len, "%s", e.what());
        return nullptr;
    }
}

void ref_destroy(void* h) { delete static_cast<Model*>(h); }

int ref_n_cols(void* h, int cold_disk) { return n_cols(*static_cast<Model*>(h), cold_disk); }

// 0: stepped; 1: RadiusCollapseException (cpp/include/application.hpp:31-42)
int ref_step(void* h) { return do_step(*static_cast<Model*>(h)); }

long ref_last_picard(void* h) { return static_cast<Model*>(h)->last_picard; }

void ref_row(void* h, int cold_disk, double* out) { fill_row(*static_cast<Model*>(h), cold_disk, out); }

void ref_state(void* h, double* t, int64_t* i_t, int64_t* first, int64_t* last) {
    FreddiEvolution* f = static_cast<Model*>(h)->ev();
    if (t) *t = f->t();
    if (i_t) *i_t = static_cast<int64_t>(f->i_t());
    if (first) *first = static_cast<int64_t>(f->first());
    if (last) *last = static_cast<int64_t>(f->last());
}

void ref_info(void* h, fb200_model_info_t* o) {
    Model& m = *static_cast<Model*>(h);
    FreddiEvolution* f = m.ev();
    std::memset(o, 0, sizeof(*o));
    o->GM = f->GM(); o->eta = f->eta(); o->R_g = f->R_g(); o->cosi = f->cosi(); o->distance = f->distance();
    o->cosiOverD2 = f->cosiOverD2(); o->tau = f->args().calc->tau;
    o->h_in = f->h().front(); o->h_out = f->h().back();
    o->rin = f->args().basic->rin; o->rout = f->args().basic->rout; o->risco = f->args().basic->risco;
    o->F0 = NAN; o->Mdisk0 = NAN; o->Mdot0 = NAN;  // protected in the reference (arguments.hpp:104-106)
    o->Nx = static_cast<int64_t>(f->Nx()); o->Nt = static_cast<int64_t>(f->Nt()); o->first0 = static_cast<int64_t>(f->first());
    if (m.is_ns) {
        auto* n = m.ns.get();
        o->mu_magn = n->mu_magn(); o->R_cor = n->R_cor(); o->R_dead = n->R_dead(); o->inverse_beta = n->inverse_beta();
        o->R_x = n->R_x(); o->redshift = n->redshift();
    }
}

// Radial fields exactly as the getters return them (no NaN padding here).
int ref_field(void* h, int field, double* out) {
    Model& m = *static_cast<Model*>(h);
    FreddiEvolution* f = m.ev();
    vecd tmp;
    const vecd* v = nullptr;
    switch (field) {
        case FB200_FIELD_H: v = &f->h(); break;
        case FB200_FIELD_R: v = &f->R(); break;
        case FB200_FIELD_F: v = &f->F(); break;
        case FB200_FIELD_W: v = &f->W(); break;
        case FB200_FIELD_SIGMA: v = &f->Sigma(); break;
        case FB200_FIELD_TPH: v = &f->Tph(); break;
        case FB200_FIELD_TPH_VIS: v = &f->Tph_vis(); break;
        case FB200_FIELD_TIRR: v = &f->Tirr(); break;
        case FB200_FIELD_KIRR: v = &f->Kirr(); break;
        case FB200_FIELD_HEIGHT: v = &f->Height(); break;
        case FB200_FIELD_TPH_X: v = &f->Tph_X(); break;
        case FB200_FIELD_WINDA: tmp = f->windA(); v = &tmp; break;
        case FB200_FIELD_WINDB: tmp = f->windB(); v = &tmp; break;
        case FB200_FIELD_WINDC: tmp = f->windC(); v = &tmp; break;
        case FB200_FIELD_FMAGN: if (!m.is_ns) return -1; v = &m.ns->Fmagn(); break;
        case FB200_FIELD_DFMAGN_DH: if (!m.is_ns) return -1; v = &m.ns->dFmagn_dh(); break;
        case FB200_FIELD_D2FMAGN_DH2: if (!m.is_ns) return -1; v = &m.ns->d2Fmagn_dh2(); break;
        default: return -1;  // Qx is protected in the reference
    }
    std::memcpy(out, v->data(), v->size() * sizeof(double));
    return 0;
}

double ref_flux(void* h, int region, double lambda) {
    FreddiEvolution* f = static_cast<Model*>(h)->ev();
    return region == 0 ? f->flux(lambda) : f->flux_region<FreddiState::ColdRegion>(lambda);
}

double ref_flux_star(void* h, double lambda, double phase) { return static_cast<Model*>(h)->ev()->flux_star(lambda, phase); }

void ref_define_mdot0(double v) { g_define_mdot0 = v; }

// Spectrum::Planck_nu1_nu2 of the reference (cpp/src/spectrum.cpp:26-37) at the tolerance FreddiState::Lx passes (1e-4,
// cpp/src/freddi_state.cpp:339-345) for n temperatures: the band integral the library's walk restates and its table tabulates
void ref_planck_nu1_nu2(const double* T, long n, double nu1, double nu2, double tol, double* out) {
    for (long i = 0; i < n; ++i) out[i] = Spectrum::Planck_nu1_nu2(T[i], nu1, nu2, tol);
}

// trapz of the reference (cpp/src/util.cpp:9-19) on caller data
double ref_trapz(const double* x, const double* y, size_t n, size_t first, size_t last) {
    return trapz(std::vector<double>(x, x + n), std::vector<double>(y, y + n), first, last);
}

double ref_mdot_wind(void* h) { return static_cast<Model*>(h)->ev()->Mdot_wind(); }

// A whole light curve the way the Python iterator produces it (cpp/include/freddi_evolution.hpp:34-39,56-57):
// Nt+1 rows, Nt steps.  Returns the number of valid rows (fewer than Nt+1 after a collapse), <0 on a
// construction error.  rows[(Nt+1) x n_cols]; Fsnap (optional) [(Nt+1) x Nx]; first/last/picard (optional) [Nt+1].
int64_t ref_run(const fb200_args_t* a, const double* lambdas, int n_lambdas, int cold_disk, double* rows, int64_t rows_cap,
                double* Fsnap, int64_t* first, int64_t* last, int64_t* picard, char* err, size_t err_len) {
    int code = 0;
    Model* m = static_cast<Model*>(ref_create(a, lambdas, n_lambdas, &code, err, err_len));
    if (!m) return code;
    const int nc = n_cols(*m, cold_disk);
    const int64_t Nt = static_cast<int64_t>(m->ev()->Nt());
    const size_t Nx = m->ev()->Nx();
    int64_t r = 0;
    for (; r <= Nt && r < rows_cap; ++r) {
        fill_row(*m, cold_disk, rows + r * nc);
        if (Fsnap) std::memcpy(Fsnap + r * Nx, m->ev()->F().data(), Nx * sizeof(double));
        if (first) first[r] = static_cast<int64_t>(m->ev()->first());
        if (last) last[r] = static_cast<int64_t>(m->ev()->last());
        if (picard) picard[r] = (r == 0) ? 0 : m->last_picard;
        if (r == Nt) { ++r; break; }
        if (do_step(*m) != 0) { ++r; break; }
    }
    delete m;
    return r;
}

// CPU baseline: n_models light curves striped over n_threads std::threads, one reference instance per
// model (instances share no mutable state).  Construction (incl. the companion-star triangulation,
// cpp/src/freddi_state.cpp:78-79) is timed separately from the dump/step loop.
// rows may be NULL (rows are still evaluated, into a scratch buffer); picard_total (optional) [n_models]: Picard iterations
// summed over the completed steps of each model.  Returns total model-steps done.
int64_t ref_run_ensemble(const fb200_args_t* args, int64_t n_models, const double* lambdas, int n_lambdas, int cold_disk,
                         int n_threads, int64_t max_steps, double* rows, int64_t rows_per_model, int32_t* rows_valid,
                         double* construct_seconds, double* evolve_seconds, int64_t* picard_total) {
    if (n_threads <= 0) n_threads = static_cast<int>(std::thread::hardware_concurrency());
    if (n_threads <= 0) n_threads = 1;
    std::vector<Model*> models(static_cast<size_t>(n_models), nullptr);
    std::atomic<int64_t> next{0};
    auto t0 = std::chrono::steady_clock::now();
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t)
            pool.emplace_back([&] {
                for (int64_t i; (i = next.fetch_add(1)) < n_models;)
                    models[static_cast<size_t>(i)] = static_cast<Model*>(ref_create(&args[i], lambdas, n_lambdas, nullptr, nullptr, 0));
            });
        for (auto& th : pool) th.join();
    }
    auto t1 = std::chrono::steady_clock::now();
    std::atomic<int64_t> steps_done{0};
    next = 0;
    {
        std::vector<std::thread> pool;
        for (int t = 0; t < n_threads; ++t)
            pool.emplace_back([&] {
                std::vector<double> scratch;
                for (int64_t i; (i = next.fetch_add(1)) < n_models;) {
                    Model* m = models[static_cast<size_t>(i)];
                    if (!m) { if (rows_valid) rows_valid[i] = 0; continue; }
                    const int nc = n_cols(*m, cold_disk);
                    scratch.resize(static_cast<size_t>(nc));
                    int64_t Nt = static_cast<int64_t>(m->ev()->Nt());
                    if (max_steps >= 0 && Nt > max_steps) Nt = max_steps;
                    int64_t r = 0, steps = 0, picard = 0;
                    for (; r <= Nt; ++r) {
                        double* dst = (rows && r < rows_per_model) ? rows + (i * rows_per_model + r) * nc : scratch.data();
                        fill_row(*m, cold_disk, dst);
                        if (r == Nt) { ++r; break; }
                        ++steps;
                        if (do_step(*m) != 0) { ++r; break; }
                        picard += m->last_picard;  // completed steps only, like the kernels' counter
                    }
                    if (rows_valid) rows_valid[i] = static_cast<int32_t>(r);
                    if (picard_total) picard_total[i] = picard;
                    steps_done += steps;
                }
            });
        for (auto& th : pool) th.join();
    }
    auto t2 = std::chrono::steady_clock::now();
    for (Model* m : models) delete m;
    if (construct_seconds) *construct_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (evolve_seconds) *evolve_seconds = std::chrono::duration<double>(t2 - t1).count();
    return steps_done.load();
}

}  // extern "C"
