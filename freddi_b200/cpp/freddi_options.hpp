// Command-line / INI front-end of the B200 ensemble engine: the options of the reference's `freddi` and
// `freddi-ns` executables, resolved to fb200_args_t (CGS) for the C ABI of include/freddi_b200.h.
//
// Stands in for (paths relative to the reference repository):
//   option tables, defaults, units     cpp/src/options.cpp:21-32,76-92,203-249,266-275,324-340,368-378
//                                      cpp/src/ns/ns_options.cpp:112-158,241-247
//   parse order                        parseOptions, cpp/include/options.hpp:91-135: command line, then --config, then
//                                      ./freddi.ini, $XDG_CONFIG_HOME|~/.config/freddi.ini, PREFIX/etc, /etc;
//                                      the first value stored wins, `lambda` / `passband` compose across sources
//   Options -> Arguments               cpp/src/options.cpp:11-18,36-74,95-201,252-264,278-322,343-366,382-401
//                                      cpp/src/ns/ns_options.cpp:3-110,161-239
// boost::program_options is not available here (and is not part of the hot path); this is a plain C++17
// restatement of the subset of its behaviour the two executables rely on: `--name=value`, `--name value`,
// unambiguous prefixes of long names, `-x value` / `-xvalue`, multitoken `--lambda a b c`, flags, positional
// tokens ignored, INI files with `#` comments and [sections].
//
// One extension, because the engine evolves ensembles: `--sweep name=lo:hi:n[:log]` (repeatable) replaces the
// scalar option `name` by n values; the cartesian product of all sweeps is evolved in one launch.  `--devices all` or
// `--devices 0,1,...` spreads the models over several GPUs of the box from this one process (model k -> device k mod G).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/freddi_b200.h"

namespace fb200 {
namespace cli {

// constants of cpp/include/gsl_const_cgsm.h and cpp/include/constants.hpp, conversions of cpp/include/unit_transformation.hpp:8-55
constexpr double kSolarMass = 1.98892e33, kSolarRadius = 6.955e10, kParsec = 3.08567758135e18, kElectronVolt = 1.602176487e-12,
                 kPlanck = 6.62606896e-27, kSpeedOfLight = 2.99792458e10, kGravConst = 6.673e-8;
inline double sunToGram(double m) { return m * kSolarMass; }
inline double rgToCm(double r, double mass_gram) { return r * kGravConst * mass_gram / (kSpeedOfLight * kSpeedOfLight); }
inline double kpcToCm(double d) { return d * 1000. * kParsec; }
inline double sunToCm(double r) { return r * kSolarRadius; }
inline double cmToSun(double r) { return r / kSolarRadius; }
inline double angstromToCm(double l) { return l * 1e-8; }
inline double cmToAngstrom(double l) { return l * 1e8; }
inline double sToDay(double t) { return t / 86400.; }
inline double dayToS(double t) { return t * 86400.; }
inline double kevToHertz(double e) { return e * 1000. * kElectronVolt / kPlanck; }
inline double hertzToKev(double nu) { return nu * kPlanck / (1000. * kElectronVolt); }

struct OptionError : std::runtime_error {  // boost::program_options::error
    using std::runtime_error::runtime_error;
};

enum class Kind { Flag, String, Double, UInt, DoubleVec, StringVec };

struct Option {
    std::string name;
    char short_name;
    Kind kind;
    bool has_default;
    std::string def;  // default value as the reference's option table gives it (text for strings, see Value::from_default)
    double def_d;
    unsigned def_u;
    bool required;
};

// One stored option value (boost::program_options::variable_value)
struct Value {
    Kind kind = Kind::Flag;
    std::string s;
    double d = 0.;
    unsigned u = 0;
    std::vector<double> dv;
    std::vector<std::string> sv;
    bool defaulted = false;
};

typedef std::map<std::string, Value> VariablesMap;  // ordered by name, like boost's variables_map (the echo relies on it)

inline Option opt_flag(const char* n, char s = 0) { return {n, s, Kind::Flag, false, "", 0., 0u, false}; }
inline Option opt_str(const char* n, char s = 0) { return {n, s, Kind::String, false, "", 0., 0u, false}; }
inline Option opt_str(const char* n, const char* def, char s = 0) { return {n, s, Kind::String, true, def, 0., 0u, false}; }
inline Option opt_dbl(const char* n, char s = 0) { return {n, s, Kind::Double, false, "", 0., 0u, false}; }
inline Option opt_dbl_def(const char* n, double def, char s = 0) { return {n, s, Kind::Double, true, "", def, 0u, false}; }
inline Option opt_dbl_req(const char* n, char s = 0) { return {n, s, Kind::Double, false, "", 0., 0u, true}; }
inline Option opt_uint(const char* n, unsigned def) { return {n, 0, Kind::UInt, true, "", 0., def, false}; }
inline Option opt_dvec(const char* n) { return {n, 0, Kind::DoubleVec, false, "", 0., 0u, false}; }
inline Option opt_svec(const char* n) { return {n, 0, Kind::StringVec, false, "", 0., 0u, false}; }

// The option table of `freddi` (FreddiOptions::description, cpp/src/options.cpp:392-401) or `freddi-ns`
// (FreddiNeutronStarOptions::description, cpp/src/ns/ns_options.cpp:249-259).
inline std::vector<Option> option_table(bool ns) {
    std::vector<Option> t = {
        // GeneralOptions, cpp/src/options.cpp:21-32
        opt_flag("help", 'h'), opt_str("config"), opt_str("prefix", "freddi"), opt_flag("stdout"), opt_str("dir", ".", 'd'),
        opt_uint("precision", 12), opt_uint("tempsparsity", 1), opt_flag("fulldata"),
        // BasicDiskBinaryOptions, cpp/src/options.cpp:76-92
        opt_dbl_req("alpha", 'a'), opt_dbl("alphacold"), opt_dbl_req("Mx", 'M'), opt_dbl_def("kerr", 0.0), opt_dbl_req("Mopt"),
        opt_dbl_def("rochelobefill", 1.0), opt_dbl_def("Topt", 0.0), opt_dbl_req("period", 'P'), opt_dbl("rin"), opt_dbl("rout", 'R'),
        opt_dbl("risco"),
        // DiskStructureOptions, cpp/src/options.cpp:203-249
        opt_str("opacity", "Kramers", 'O'), opt_dbl_def("Mdotout", 0.), opt_str("boundcond", "Teff"), opt_dbl_def("Thot", 0.),
        opt_dbl_def("Qirr2Qvishot", 0.), opt_str("initialcond", "powerF"), opt_dbl("F0"), opt_dbl("Mdisk0"), opt_dbl("Mdot0"),
        opt_dbl("powerorder"), opt_dbl("gaussmu"), opt_dbl("gausssigma"), opt_str("windtype", "no"), opt_dbl("windC_w"),
        opt_dbl("windR_w"), opt_dbl("windA_0"), opt_dbl("windB_1"), opt_dbl("windXi_max"), opt_dbl("windT_ic"), opt_dbl("windPow"),
        opt_dbl("windC_0"),
    };
    if (ns) {  // NeutronStarOptions, cpp/src/ns/ns_options.cpp:112-158
        const double third = 1.0 / 3.0;
        const std::vector<Option> n = {
            opt_str("nsprop", "dummy"), opt_dbl("freqx"), opt_dbl("Rx"), opt_dbl_req("Bx"), opt_dbl_def("hotspotarea", 1.),
            opt_dbl_def("epsilonAlfven", 1.), opt_dbl_def("inversebeta", 0.), opt_dbl_def("Rdead", 0.), opt_str("fptype", "no-outflow"),
            opt_dbl("fp-geometrical-chi"), opt_dbl("romanova2018-par1"), opt_dbl("romanova2018-par2"), opt_str("kappattype", "const"),
            opt_dbl_def("kappat-const-value", third), opt_dbl_def("kappat-corstep-in", third), opt_dbl_def("kappat-corstep-out", third),
            opt_dbl_def("kappat-romanova2018-in", third), opt_dbl_def("kappat-romanova2018-out", third), opt_str("nsgravredshift", "off"),
        };
        t.insert(t.end(), n.begin(), n.end());
    }
    const std::vector<Option> rest = {
        // SelfIrradiationOptions, cpp/src/options.cpp:266-275 (+ ns_options.cpp:241-247)
        opt_dbl_def("Cirr", 0.), opt_dbl_def("irrindex", 0.), opt_dbl_def("Cirrcold", 0.), opt_dbl_def("irrindexcold", 0.),
        opt_dbl_def("h2rcold", 0.), opt_str("angulardistdisk", "plane"),
        // FluxOptions, cpp/src/options.cpp:324-340
        opt_dbl_def("colourfactor", 1.7), opt_dbl_def("emin", hertzToKev(kevToHertz(1.))), opt_dbl_def("emax", hertzToKev(kevToHertz(12.))),
        opt_dbl_def("staralbedo", 0.), opt_dbl_def("inclination", 0., 'i'), opt_dbl_def("ephemerist0", 0.), opt_dbl_req("distance"),
        opt_flag("colddiskflux"), opt_flag("starflux"), opt_dvec("lambda"), opt_svec("passband"),
        // CalculationOptions, cpp/src/options.cpp:368-378
        opt_dbl_def("inittime", 0.), opt_dbl_req("time", 'T'), opt_dbl("tau"), opt_uint("Nx", 1000), opt_str("gridscale", "log"),
        opt_uint("starlod", 3),
        // extension of this front-end (not echoed into the output files)
        opt_svec("sweep"), opt_str("devices"),
    };
    t.insert(t.end(), rest.begin(), rest.end());
    if (ns) t.push_back(opt_str("angulardistns", "isotropic"));
    return t;
}

class Parser {
public:
    explicit Parser(bool ns) : table_(option_table(ns)) {}

    const Option* find_long(const std::string& name, bool guess) const {
        const Option* prefix_hit = nullptr;
        int hits = 0;
        for (const Option& o : table_) {
            if (o.name == name) return &o;
            if (guess && o.name.compare(0, name.size(), name) == 0) { prefix_hit = &o; ++hits; }
        }
        if (hits == 1) return prefix_hit;
        if (hits > 1) throw OptionError("option '--" + name + "' is ambiguous");
        return nullptr;
    }
    const Option* find_short(char c) const {
        for (const Option& o : table_) if (o.short_name == c) return &o;
        return nullptr;
    }

    // One source (the command line or one file): name -> raw tokens in order of appearance
    typedef std::vector<std::pair<const Option*, std::vector<std::string>>> Parsed;

    Parsed parse_command_line(int ac, const char* const* av) const {
        Parsed out;
        for (int i = 1; i < ac; ++i) {
            const std::string tok = av[i];
            const Option* o = nullptr;
            std::vector<std::string> vals;
            bool adjacent = false;
            if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-') {
                const size_t eq = tok.find('=');
                const std::string name = tok.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
                o = find_long(name, true);
                if (!o) throw OptionError("unrecognised option '--" + name + "'");
                if (eq != std::string::npos) { vals.push_back(tok.substr(eq + 1)); adjacent = true; }
            } else if (tok.size() >= 2 && tok[0] == '-' && !is_number(tok)) {
                o = find_short(tok[1]);
                if (!o) throw OptionError(std::string("unrecognised option '-") + tok[1] + "'");
                if (tok.size() > 2) { vals.push_back(tok.substr(2)); adjacent = true; }
            } else {
                continue;  // a positional token: program_options stores it without a name and store() skips it
            }
            if (o->kind == Kind::Flag) {
                if (adjacent) throw OptionError("option '--" + o->name + "' does not take any arguments");
            } else if (o->kind == Kind::DoubleVec || o->kind == Kind::StringVec) {
                while (i + 1 < ac && !looks_like_option(av[i + 1])) vals.push_back(av[++i]);  // multitoken()
                if (vals.empty()) throw OptionError("the required argument for option '--" + o->name + "' is missing");
            } else if (!adjacent) {
                if (i + 1 >= ac) throw OptionError("the required argument for option '--" + o->name + "' is missing");
                vals.push_back(av[++i]);
            }
            out.emplace_back(o, vals);
        }
        return out;
    }

    // boost::program_options::parse_config_file: '#' starts a comment, [section] prefixes names with "section.",
    // every other non-empty line is name=value; unknown names are errors
    Parsed parse_config_file(std::istream& in) const {
        Parsed out;
        std::string line, section;
        while (std::getline(in, line)) {
            const size_t hash = line.find('#');
            if (hash != std::string::npos) line.erase(hash);
            line = trim(line);
            if (line.empty()) continue;
            if (line.front() == '[' && line.back() == ']') {
                section = line.substr(1, line.size() - 2);
                if (!section.empty() && section.back() != '.') section += '.';
                continue;
            }
            const size_t eq = line.find('=');
            if (eq == std::string::npos) throw OptionError("the options configuration file contains an invalid line '" + line + "'");
            const std::string name = section + trim(line.substr(0, eq));
            const Option* o = find_long(name, false);
            if (!o) throw OptionError("unrecognised option '" + name + "'");
            out.emplace_back(o, std::vector<std::string>{trim(line.substr(eq + 1))});
        }
        return out;
    }

    // boost::program_options::store: a value already stored (by an earlier source) is kept; composing options
    // (lambda, passband, sweep) collect the values of every source; an option given twice in one source is an error
    void store(const Parsed& parsed, VariablesMap& vm) const {
        std::map<std::string, bool> new_final;
        for (const auto& pv : parsed) {
            const Option& o = *pv.first;
            const bool composing = (o.kind == Kind::DoubleVec || o.kind == Kind::StringVec);
            if (final_.count(o.name)) continue;  // stored by an earlier source
            if (!composing && new_final.count(o.name)) throw OptionError("option '--" + o.name + "' cannot be specified more than once");
            Value& v = vm[o.name];
            v.kind = o.kind;
            for (const std::string& tokn : pv.second) {
                switch (o.kind) {
                    case Kind::Flag: v.s = ""; break;
                    case Kind::String: v.s = tokn; break;
                    case Kind::Double: v.d = to_double(tokn, o.name); break;
                    case Kind::UInt: v.u = to_uint(tokn, o.name); break;
                    case Kind::DoubleVec: v.dv.push_back(to_double(tokn, o.name)); break;
                    case Kind::StringVec: v.sv.push_back(tokn); break;
                }
            }
            if (!composing) new_final[o.name] = true;
        }
        final_.insert(new_final.begin(), new_final.end());
    }

    // defaults + boost::program_options::notify (required options)
    void finish(VariablesMap& vm) const {
        for (const Option& o : table_) {
            if (vm.count(o.name) || !o.has_default) continue;
            Value v;
            v.kind = o.kind;
            v.defaulted = true;
            v.s = o.def; v.d = o.def_d; v.u = o.def_u;
            vm[o.name] = v;
        }
        for (const Option& o : table_)
            if (o.required && !vm.count(o.name)) throw OptionError("the option '--" + o.name + "' is required but missing");
    }

    const std::vector<Option>& table() const { return table_; }

    static std::string trim(const std::string& s) {
        const char* ws = " \t\r\n";
        const size_t b = s.find_first_not_of(ws);
        if (b == std::string::npos) return "";
        return s.substr(b, s.find_last_not_of(ws) - b + 1);
    }
    static bool is_number(const std::string& s) {
        char* end = nullptr;
        std::strtod(s.c_str(), &end);
        return end && *end == 0 && end != s.c_str();
    }
    static bool looks_like_option(const std::string& s) { return s.size() >= 2 && s[0] == '-' && !is_number(s); }
    static double to_double(const std::string& s, const std::string& name) {  // boost::lexical_cast<double>: the whole token
        char* end = nullptr;
        const double v = std::strtod(s.c_str(), &end);
        if (s.empty() || !end || *end != 0 || std::isspace(static_cast<unsigned char>(s[0])))
            throw OptionError("the argument ('" + s + "') for option '--" + name + "' is invalid");
        return v;
    }
    static unsigned to_uint(const std::string& s, const std::string& name) {
        char* end = nullptr;
        if (s.empty() || s[0] == '-' || s[0] == '+' || std::isspace(static_cast<unsigned char>(s[0])))
            throw OptionError("the argument ('" + s + "') for option '--" + name + "' is invalid");
        const unsigned long v = std::strtoul(s.c_str(), &end, 10);
        if (!end || *end != 0 || v > 0xfffffffful) throw OptionError("the argument ('" + s + "') for option '--" + name + "' is invalid");
        return static_cast<unsigned>(v);
    }

private:
    std::vector<Option> table_;
    mutable std::map<std::string, bool> final_;  // boost: variables_map::m_final
};

// parseOptions (cpp/include/options.hpp:91-135).  Returns false when --help was given (usage goes to `help_out`).
inline bool parse_options(bool ns, int ac, const char* const* av, VariablesMap& vm, std::string* help_out = nullptr,
                          const std::string& install_prefix = "/usr/local") {
    Parser p(ns);
    p.store(p.parse_command_line(ac, av), vm);
    if (vm.count("help")) {
        if (help_out) {
            std::ostringstream os;
            os << (ns ? "Freddi NS" : "Freddi") << ": numerical calculation of accretion disk evolution (B200 ensemble engine)\n";
            for (const Option& o : p.table()) {
                os << "  ";
                if (o.short_name) os << '-' << o.short_name << " [ --" << o.name << " ]"; else os << "--" << o.name;
                if (o.kind != Kind::Flag) os << " arg";
                if (o.has_default) {
                    if (o.kind == Kind::String) os << " (=" << o.def << ")";
                    else if (o.kind == Kind::Double) os << " (=" << o.def_d << ")";
                    else if (o.kind == Kind::UInt) os << " (=" << o.def_u << ")";
                }
                os << "\n";
            }
            *help_out = os.str();
        }
        return false;
    }
    std::vector<std::string> paths;
    const char* xdg = std::getenv("XDG_CONFIG_HOME");
    const char* home = std::getenv("HOME");
    const std::string config_home = xdg ? std::string(xdg) : std::string(home ? home : "") + "/.config";
    for (const std::string& dir : {std::string("."), config_home, install_prefix + "/etc", std::string("/etc")}) paths.push_back(dir + "/freddi.ini");
    if (vm.count("config")) paths.insert(paths.begin(), vm["config"].s);
    for (const std::string& path : paths) {
        std::ifstream f(path);
        if (!f) continue;
        p.store(p.parse_config_file(f), vm);
    }
    p.finish(vm);
    return true;
}

// ---------------------------------------------------------------------------------------------------------------
// Options -> Arguments
// ---------------------------------------------------------------------------------------------------------------
inline int enum_of(const std::string& v, std::initializer_list<const char*> names) {
    int k = 0;
    for (const char* n : names) { if (v == n) return k; ++k; }
    return 0x7fff;  // out of range: the resolver raises the error class the reference's constructor does
}

inline double dbl(const VariablesMap& vm, const char* name) { return vm.at(name).d; }
inline double opt_dbl_or_nan(const VariablesMap& vm, const char* name) { return vm.count(name) ? vm.at(name).d : NAN; }  // varToOpt
inline void need(const VariablesMap& vm, const char* name, const char* what, const std::string& which) {
    if (!vm.count(name)) throw OptionError(std::string("--") + name + " is required if --" + what + "=" + which);
}

struct RunSetup {
    fb200_args_t args;
    // GeneralArguments (cpp/include/arguments.hpp:18-44)
    std::string prefix, dir;
    unsigned precision = 12, tempsparsity = 1;
    bool fulldata = false, to_stdout = false;
    // FluxArguments' ensemble-wide parts
    bool cold_disk = false, star_flux = false;
    std::vector<double> lambdas;            // cm
    std::vector<std::string> passband_files;
};

inline RunSetup make_setup(const VariablesMap& vm, bool ns) {
    RunSetup s;
    fb200_args_t& a = s.args;
    fb200_args_default(&a, ns ? 1 : 0);
    // FreddiOptions ctor, cpp/src/options.cpp:383-386
    if (vm.count("starflux") && (!vm.count("Mx") || !vm.count("Topt") || !vm.count("Mopt") || !vm.count("period")))
        throw OptionError("--starflux requires --Mx, --Mopt and --period to be specified");
    s.star_flux = vm.count("starflux") > 0;
    // GeneralOptions, cpp/src/options.cpp:11-18
    s.prefix = vm.at("prefix").s; s.dir = vm.at("dir").s; s.precision = vm.at("precision").u; s.tempsparsity = vm.at("tempsparsity").u;
    s.fulldata = vm.count("fulldata") > 0; s.to_stdout = vm.count("stdout") > 0;
    if (s.tempsparsity == 0) throw OptionError("--tempsparsity must be positive");
    // BasicDiskBinaryOptions, cpp/src/options.cpp:36-74
    a.alpha = dbl(vm, "alpha"); a.alphacold = opt_dbl_or_nan(vm, "alphacold");
    a.Mx = sunToGram(dbl(vm, "Mx")); a.kerr = dbl(vm, "kerr"); a.period = dayToS(dbl(vm, "period")); a.Mopt = sunToGram(dbl(vm, "Mopt"));
    a.rochelobefill = dbl(vm, "rochelobefill"); a.Topt = dbl(vm, "Topt");
    a.rin = vm.count("rin") ? rgToCm(dbl(vm, "rin"), a.Mx) : NAN;
    a.rout = vm.count("rout") ? sunToCm(dbl(vm, "rout")) : NAN;
    a.risco = vm.count("risco") ? rgToCm(dbl(vm, "risco"), a.Mx) : NAN;
    // DiskStructureOptions, cpp/src/options.cpp:95-201
    a.opacity = enum_of(vm.at("opacity").s, {"Kramers", "OPAL"});
    a.Mdotout = dbl(vm, "Mdotout");
    a.boundcond = enum_of(vm.at("boundcond").s, {"Teff", "Tirr"});
    a.Thot = dbl(vm, "Thot");
    a.Qirr2Qvishot = dbl(vm, "Qirr2Qvishot");  // the C ABI takes the ratio and stores its 4th root, as options.cpp:103 does
    a.initialcond = enum_of(vm.at("initialcond").s, {"powerF", "linearF", "powerSigma", "sineF", "gaussF", "quasistat", "quasistat-ns"});
    if (!ns && a.initialcond == FB200_IC_QUASISTAT_NS) a.initialcond = 0x7fff;  // only freddi-ns knows it
    {   // aliases accepted by DiskStructureArguments::initializeInitialFFunction (cpp/src/arguments.cpp:70,98,194)
        const std::string& ic = vm.at("initialcond").s;
        if (ic == "power") a.initialcond = FB200_IC_POWERF;
        if (ic == "linear") a.initialcond = FB200_IC_LINEARF;
        if (ic == "sinusF" || ic == "sinus") a.initialcond = FB200_IC_SINEF;
    }
    a.F0 = opt_dbl_or_nan(vm, "F0"); a.Mdisk0 = opt_dbl_or_nan(vm, "Mdisk0"); a.Mdot0 = opt_dbl_or_nan(vm, "Mdot0");
    a.powerorder = opt_dbl_or_nan(vm, "powerorder"); a.gaussmu = opt_dbl_or_nan(vm, "gaussmu"); a.gausssigma = opt_dbl_or_nan(vm, "gausssigma");
    {
        const std::string& w = vm.at("windtype").s;  // windparamsInitializer, cpp/src/options.cpp:114-201
        double* p = a.windparams;
        if (w == "no") a.windtype = FB200_WIND_NO;
        else if (w == "SS73C") a.windtype = FB200_WIND_SS73C;
        else if (w == "ShieldsOscil1986") { need(vm, "windC_w", "windtype", w); need(vm, "windR_w", "windtype", w); a.windtype = FB200_WIND_SHIELDSOSCIL1986; p[0] = dbl(vm, "windC_w"); p[1] = dbl(vm, "windR_w"); }
        else if (w == "Janiuk2015") { need(vm, "windA_0", "windtype", w); need(vm, "windB_1", "windtype", w); a.windtype = FB200_WIND_JANIUK2015; p[0] = dbl(vm, "windA_0"); p[1] = dbl(vm, "windB_1"); }
        else if (w == "Shields1986" || w == "Woods1996") {
            need(vm, "windXi_max", "windtype", w); need(vm, "windT_ic", "windtype", w); need(vm, "windPow", "windtype", w);
            a.windtype = (w == "Shields1986") ? FB200_WIND_SHIELDS1986 : FB200_WIND_WOODS1996;
            p[0] = dbl(vm, "windXi_max"); p[1] = dbl(vm, "windT_ic"); p[2] = dbl(vm, "windPow");
        }
        else if (w == "Woods1996AGN") { need(vm, "windC_0", "windtype", w); need(vm, "windT_ic", "windtype", w); a.windtype = FB200_WIND_WOODS1996AGN; p[0] = dbl(vm, "windC_0"); p[1] = dbl(vm, "windT_ic"); }
        else if (w == "toy") { need(vm, "windC_w", "windtype", w); a.windtype = FB200_WIND_TOY; p[0] = dbl(vm, "windC_w"); }
        else throw OptionError("Unknown --windtype=" + w);
    }
    // SelfIrradiationOptions, cpp/src/options.cpp:252-264
    a.Cirr = dbl(vm, "Cirr"); a.irrindex = dbl(vm, "irrindex"); a.Cirrcold = dbl(vm, "Cirrcold"); a.irrindexcold = dbl(vm, "irrindexcold");
    a.h2rcold = dbl(vm, "h2rcold");
    a.angulardistdisk = enum_of(vm.at("angulardistdisk").s, {"plane", "isotropic"});
    if (a.Cirr <= 0. && a.boundcond == FB200_BOUNDCOND_TIRR) throw OptionError("Set positive --Cirr when --boundcond=Tirr");
    // FluxOptions, cpp/src/options.cpp:278-322
    a.colourfactor = dbl(vm, "colourfactor"); a.emin = kevToHertz(dbl(vm, "emin")); a.emax = kevToHertz(dbl(vm, "emax"));
    a.staralbedo = dbl(vm, "staralbedo"); a.inclination = dbl(vm, "inclination"); a.ephemerist0 = dayToS(dbl(vm, "ephemerist0"));
    a.distance = kpcToCm(dbl(vm, "distance"));
    s.cold_disk = vm.count("colddiskflux") > 0;
    if (vm.count("lambda")) for (double l : vm.at("lambda").dv) s.lambdas.push_back(angstromToCm(l));
    if (vm.count("passband")) s.passband_files = vm.at("passband").sv;
    // CalculationOptions, cpp/src/options.cpp:343-366
    a.inittime = dayToS(dbl(vm, "inittime")); a.time = dayToS(dbl(vm, "time")); a.tau = vm.count("tau") ? dayToS(dbl(vm, "tau")) : NAN;
    a.Nx = vm.at("Nx").u;
    a.gridscale = enum_of(vm.at("gridscale").s, {"log", "linear"});
    if (a.gridscale == 0x7fff) throw OptionError("Invalid --gridscale value");
    a.starlod = int(vm.at("starlod").u);
    if (ns) {  // NeutronStarOptions, cpp/src/ns/ns_options.cpp:3-110
        a.nsprop = enum_of(vm.at("nsprop").s, {"dummy", "sibgatullinsunyaev2000", "newt"});
        a.freqx = opt_dbl_or_nan(vm, "freqx"); a.Rx = opt_dbl_or_nan(vm, "Rx"); a.Bx = dbl(vm, "Bx");
        a.hotspotarea = dbl(vm, "hotspotarea"); a.epsilonAlfven = dbl(vm, "epsilonAlfven"); a.inversebeta = dbl(vm, "inversebeta");
        a.Rdead = dbl(vm, "Rdead");
        const std::string& fp = vm.at("fptype").s;
        if (fp == "no-outflow") a.fptype = FB200_FP_NO_OUTFLOW;
        else if (fp == "propeller") a.fptype = FB200_FP_PROPELLER;
        else if (fp == "corotation-block") a.fptype = FB200_FP_COROTATION_BLOCK;
        else if (fp == "geometrical") { need(vm, "fp-geometrical-chi", "fptype", fp); a.fptype = FB200_FP_GEOMETRICAL; a.fpparams[0] = dbl(vm, "fp-geometrical-chi"); }
        else if (fp == "eksi-kultu2010") a.fptype = 0x7fff;  // the CLI's spelling is unknown to the engine (ns_options.cpp:37 vs ns_evolution.cpp:354): invalid_argument there
        else if (fp == "romanova2018") {
            need(vm, "romanova2018-par1", "fptype", fp); need(vm, "romanova2018-par2", "fptype", fp);
            a.fptype = FB200_FP_ROMANOVA2018; a.fpparams[0] = dbl(vm, "romanova2018-par1"); a.fpparams[1] = dbl(vm, "romanova2018-par2");
        } else throw OptionError("Unknown --fptype=" + fp);
        const std::string& kt = vm.at("kappattype").s;
        if (kt == "const") { a.kappattype = FB200_KAPPAT_CONST; a.kappatparams[0] = dbl(vm, "kappat-const-value"); }
        else if (kt == "corstep") { a.kappattype = FB200_KAPPAT_CORSTEP; a.kappatparams[0] = dbl(vm, "kappat-corstep-in"); a.kappatparams[1] = dbl(vm, "kappat-corstep-out"); }
        else if (kt == "romanova2018") {
            need(vm, "romanova2018-par1", "kappattype", kt); need(vm, "romanova2018-par2", "kappattype", kt);
            a.kappattype = FB200_KAPPAT_ROMANOVA2018;
            a.kappatparams[0] = dbl(vm, "kappat-romanova2018-in"); a.kappatparams[1] = dbl(vm, "kappat-romanova2018-out");
            a.kappatparams[2] = dbl(vm, "romanova2018-par1"); a.kappatparams[3] = dbl(vm, "romanova2018-par2");
        } else throw OptionError("Unknown --kappattype=" + kt);
        a.nsgravredshift = enum_of(vm.at("nsgravredshift").s, {"off", "on"});
        a.angulardistns = enum_of(vm.at("angulardistns").s, {"plane", "isotropic"});
    }
    return s;
}

// ---------------------------------------------------------------------------------------------------------------
// --sweep name=lo:hi:n[:log]
// ---------------------------------------------------------------------------------------------------------------
struct Sweep {
    std::string name;
    std::vector<double> values;
};

inline std::vector<Sweep> parse_sweeps(const VariablesMap& vm, const Parser& p) {
    std::vector<Sweep> out;
    if (!vm.count("sweep")) return out;
    for (const std::string& spec : vm.at("sweep").sv) {
        const size_t eq = spec.find('=');
        if (eq == std::string::npos) throw OptionError("--sweep expects name=lo:hi:n[:log], got '" + spec + "'");
        Sweep s;
        s.name = spec.substr(0, eq);
        const Option* o = p.find_long(s.name, false);
        if (!o || o->kind != Kind::Double) throw OptionError("--sweep: '" + s.name + "' is not a scalar floating-point option");
        std::vector<std::string> parts;
        std::stringstream ss(spec.substr(eq + 1));
        for (std::string t; std::getline(ss, t, ':');) parts.push_back(t);
        if (parts.size() < 3 || parts.size() > 4) throw OptionError("--sweep expects name=lo:hi:n[:log], got '" + spec + "'");
        const double lo = Parser::to_double(parts[0], "sweep"), hi = Parser::to_double(parts[1], "sweep");
        const unsigned n = Parser::to_uint(parts[2], "sweep");
        const bool lg = parts.size() == 4 && parts[3] == "log";
        if (parts.size() == 4 && !lg && parts[3] != "lin") throw OptionError("--sweep scale must be 'log' or 'lin'");
        if (n == 0 || (lg && !(lo > 0. && hi > 0.))) throw OptionError("--sweep: bad range in '" + spec + "'");
        for (unsigned k = 0; k < n; ++k) {
            const double f = n > 1 ? double(k) / double(n - 1) : 0.;
            s.values.push_back(lg ? lo * std::pow(hi / lo, f) : lo + (hi - lo) * f);
        }
        out.push_back(s);
    }
    return out;
}

}  // namespace cli
}  // namespace fb200
