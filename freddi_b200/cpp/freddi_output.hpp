// freddi.dat writer of the B200 ensemble engine: byte-compatible with the reference's FreddiFileOutput /
// FreddiNeutronStarFileOutput for the columns the hot path produces.
//
// Stands in for (paths relative to the reference repository):
//   outputHeader                         cpp/src/output.cpp:19-36
//   BasicFreddiFileOutput ctor           cpp/src/output.cpp:47-125  (header, "### Parameters" echo of the variables map in
//                                                                    name order, "### Derived values")
//   shortDump / diskStructureDump        cpp/src/output.cpp:134-162 (tab-separated, stream precision = --precision)
//   field names, units, descriptions     cpp/src/output.cpp:197-305, cpp/src/ns/ns_output.cpp:6-27
// Numbers are printed as a C++ ostream prints a double with precision(p) and default float format, i.e. "%.{p}g".
// The star dump of --fulldata --starflux (PREFIX_<i_t>_star.dat, cpp/src/output.cpp:164-181) is not produced.
#pragma once
#include <cstdio>
#include <ostream>
#include <string>
#include <vector>

#include "freddi_options.hpp"

namespace fb200 {
namespace cli {

struct FieldMeta {
    std::string name, unit, description;
};

inline std::string format_double(double v, unsigned precision) {
    char buf[64];
    std::snprintf(buf, sizeof buf, "%.*g", int(precision), v);
    return buf;
}

// FreddiFileOutput::initializeShortFields (cpp/src/output.cpp:197-292) + FreddiNeutronStarFileOutput (ns_output.cpp:6-19)
inline std::vector<FieldMeta> short_fields(bool ns, bool cold_disk, const std::vector<double>& lambdas_cm,
                                           const std::vector<std::string>& passband_names, bool star = false) {
    std::vector<FieldMeta> f = {
        {"t", "days", "Time moment"},
        {"Mdot", "g/s", "Accretion rate onto central object"},
        {"Mdisk", "g", "Mass of the hot disk"},
        {"Rhot", "Rsun", "Radius of the hot disk"},
        {"Sigmaout", "g/cm^2", "Surface density at the outer radius of the hot disk"},
        {"Kirrout", "float", "Irradiation coefficient Kirr at the outer radius of the hot disk"},
        {"H2R", "float", "Relative semiheight at the outer radius of the hot disk"},
        {"Teffout", "K", "Effective tempreture at the outer radius of the hot disk"},
        {"Tirrout", "K", "Irradiation temperature (Qirr / sigma_SB)^1/4 at the outer radius of the hot disk"},
        {"Qirr2Qvisout", "float", "Irradiation flux to viscous flux ratio at the outer radius of the hot disk"},
        {"TphXmax", "keV", "Maximum effective temperature of the disk"},
        {"Lx", "erg/s", "X-ray luminosity of the disk in the given energy range [emin, emax]"},
        {"Lbol", "erg/s", "Bolometric luminosity of the disk"},
        {"Fx", "erg/s/cm^2", "X-ray flux of the disk in the given energy range [emin, emax]"},
        {"Fbol", "erg/s/cm^2", "Bolometric flux of the disk"},
    };
    for (size_t i = 0; i < lambdas_cm.size(); ++i) {
        const std::string aa = std::to_string(cmToAngstrom(lambdas_cm[i])) + " AA";
        f.push_back({"Fnu" + std::to_string(i), "erg/s/cm^2/Hz", "Spectral flux density of the hot disk at wavelength of " + aa});
        if (cold_disk)
            f.push_back({"Fnu" + std::to_string(i) + "_cold", "erg/s/cm^2/Hz", "Spectral flux density of the cold disk at wavelength of " + aa});
        if (star) {
            const std::string n = "Fnu" + std::to_string(i), d = "Spectral flux density of the optical star at wavelength of " + aa;
            f.push_back({n + "_star", "erg/s/cm^2/Hz", d + " with respect to an orbital phase"});
            f.push_back({n + "_star_min", "erg/s/cm^2/Hz", d + " on the phase of inferior conjunction of the star"});
            f.push_back({n + "_star_max", "erg/s/cm^2/Hz", d + " on the phase of superior conjunction of the star"});
        }
    }
    for (const std::string& pb : passband_names) {
        f.push_back({"Fnu" + pb, "erg/s/cm^2/Hz", "Spectral flux density of the hot disk in passband " + pb});
        // unit and description are swapped in the reference (cpp/src/output.cpp:266-267); kept for byte compatibility
        if (cold_disk) f.push_back({"Fnu" + pb + "_cold", "Spectral flux density of the cold disk in passband " + pb, "erg/s/cm^2/Hz"});
        if (star) {
            const std::string d = "Spectral flux density of the optical star in passband " + pb;
            f.push_back({"Fnu" + pb + "_star", "erg/s/cm^2/Hz", d + " with respect to an orbital phase"});
            f.push_back({"Fnu" + pb + "_star_min", "erg/s/cm^2/Hz", d + " on the phase of inferior conjunction of the star"});
            f.push_back({"Fnu" + pb + "_star_max", "erg/s/cm^2/Hz", d + " on the phase of superior conjunction of the star"});
        }
    }
    if (ns) {
        const std::vector<FieldMeta> n = {
            {"Rin", "cm", "Inner radius of the disk"},
            {"etans", "float", "Accretion efficiency of the neutron star"},
            {"Lxns", "erg/s", "X-ray luminosity of the neutron star in the given energy range [emin, emax]"},
            {"Lbolns", "erg/s", "Bolometric luminosity of the neutron star"},
            {"Fxns", "erg/s/cm^2", "X-ray flux of the neutron star in the given energy range [emin, emax]"},
            {"Fbolns", "erg/s/cm^2", "Bolometric flux of the neutron star"},
            {"Thotspot", "keV", "Temperature of the neutron star 'hot spot'"},
            {"fpin", "float", "Part of accreting matter falling onto the neutron star"},
            {"Fmagnin", "dyn*cm", "Magnetic torque at the inner radius of the disk"},
            {"Fin", "dyn*cm", "Viscous torque at the inner radius of the disk"},
        };
        f.insert(f.end(), n.begin(), n.end());
    }
    return f;
}

// initializeDiskStructureFields (cpp/src/output.cpp:294-305, ns_output.cpp:21-25) with the FB200_FIELD_* each one reads
struct LongField {
    FieldMeta meta;
    int field;
};
inline std::vector<LongField> disk_structure_fields(bool ns) {
    std::vector<LongField> f = {
        {{"h", "cm^2/s", "Keplerian specific angular momentum"}, FB200_FIELD_H},
        {{"R", "cm", "Radius"}, FB200_FIELD_R},
        {{"F", "dyn*cm", "Viscous torque"}, FB200_FIELD_F},
        {{"Sigma", "g/cm^2", "Surface density"}, FB200_FIELD_SIGMA},
        {{"Teff", "K", "Effective temperature"}, FB200_FIELD_TPH},
        {{"Tvis", "K", "Viscous temperature (Qvis / sigma_SB)^1/4"}, FB200_FIELD_TPH_VIS},
        {{"Tirr", "K", "Irradiation temperature (Qirr / sigma_SB)^1/4"}, FB200_FIELD_TIRR},
        {{"Height", "cm", "Semiheight"}, FB200_FIELD_HEIGHT},
    };
    if (ns) f.push_back({{"Fmagn", "dyn*cm", "Magnetic torque"}, FB200_FIELD_FMAGN});
    return f;
}

template <class Fields, class GetMeta>
void output_header(std::ostream& os, const Fields& fields, GetMeta meta) {  // outputHeader, cpp/src/output.cpp:19-36
    os << "#" << meta(fields.at(0)).name;
    for (size_t i = 1; i < fields.size(); ++i) os << "\t" << meta(fields[i]).name;
    os << "\n";
    os << "#" << meta(fields.at(0)).unit;
    for (size_t i = 1; i < fields.size(); ++i) os << "\t" << meta(fields[i]).unit;
    os << "\n";
    os << "### Columns description\n";
    for (size_t i = 0; i < fields.size(); ++i) {
        const FieldMeta& f = meta(fields[i]);
        os << "# " << i + 1 << "=" << f.name << " [" << f.unit << "] : " << f.description << "\n";
    }
}

// "### Parameters": every stored option in name order, value printed by type (cpp/src/output.cpp:66-115)
inline void echo_parameters(std::ostream& os, const VariablesMap& vm, unsigned precision) {
    os << "### Parameters\n";
    for (const auto& kv : vm) {
        const std::string& name = kv.first;
        const Value& v = kv.second;
        if (name == "sweep" || name == "devices") continue;  // this front-end's extension; each file echoes the swept option's own value
        switch (v.kind) {
            case Kind::Flag:
            case Kind::String: os << "# " << name << "=" << v.s << "\n"; break;
            case Kind::Double: os << "# " << name << "=" << format_double(v.d, precision) << "\n"; break;
            case Kind::UInt: os << "# " << name << "=" << v.u << "\n"; break;
            case Kind::DoubleVec:
                for (size_t i = 0; i < v.dv.size(); ++i) os << "# " << name << "=" << format_double(v.dv[i], precision) << "  # " << i << "\n";
                break;
            case Kind::StringVec:
                for (size_t i = 0; i < v.sv.size(); ++i) os << "# " << name << "=" << v.sv[i] << "  # " << i << "\n";
                break;
        }
    }
}

// Everything BasicFreddiFileOutput's constructor writes (cpp/src/output.cpp:62-125)
inline void write_dat_header(std::ostream& os, const std::vector<FieldMeta>& fields, const VariablesMap& vm, unsigned precision,
                             double alphacold, double rout_cm, double risco_cm, double tau_s) {
    output_header(os, fields, [](const FieldMeta& f) -> const FieldMeta& { return f; });
    echo_parameters(os, vm, precision);
    os << "### Derived values\n"
       << "# alpha_cold = " << format_double(alphacold, precision) << "\n"
       << "# Tidal radius = " << format_double(cmToSun(rout_cm), precision) << " Rsun\n"
       << "# ISCO radius = " << format_double(risco_cm, precision) << " cm\n"
       << "# tau = " << format_double(sToDay(tau_s), precision) << " days\n";
    os.flush();
}

inline void write_row(std::ostream& os, const double* row, size_t ncol, unsigned precision) {  // shortDump, cpp/src/output.cpp:134-140
    os << format_double(row[0], precision);
    for (size_t i = 1; i < ncol; ++i) os << "\t" << format_double(row[i], precision);
    os << std::endl;
}

// diskStructureDump (cpp/src/output.cpp:142-162): columns[j] points at field j of this model (Nx values)
inline void write_disk_structure(std::ostream& os, const std::vector<LongField>& fields, const std::vector<const double*>& columns,
                                 double t_days, long first, long last, unsigned precision) {
    output_header(os, fields, [](const LongField& f) -> const FieldMeta& { return f.meta; });
    os << "### t = " << format_double(t_days, precision) << " days" << std::endl;
    for (long i = first; i <= last; ++i) {
        os << format_double(columns.at(0)[i], precision);
        for (size_t j = 1; j < fields.size(); ++j) os << "\t" << format_double(columns[j][i], precision);
        os << "\n";
    }
    os.flush();
}

}  // namespace cli
}  // namespace fb200
