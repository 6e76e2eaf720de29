// Host-side argument resolution for the ensemble engine: fb200_args_t -> what the kernels read.
// Stands in for the constructors the reference runs once per model
// (cpp/src/arguments.cpp:30-351, cpp/src/ns/ns_arguments.cpp:17-189, cpp/src/freddi_state.cpp:13-122,387-450,
//  cpp/src/ns/ns_evolution.cpp:45-139,322-383).  Runs on the host in IEEE double with the same libm the
// reference links, in the reference's operation order, because it defines F(t=0), the grid and every
// per-model constant.
#pragma once
#include <string>
#include <vector>

#include "../../include/freddi_b200.h"
#include "fb200_model.h"

namespace fb200 {

struct Resolved {
    fb200_model_dev_t dev;
    fb200_model_info_t info;
    std::vector<double> h, F;              // [Nx]
    std::vector<double> wA, wB, wC;        // static wind arrays [Nx]; empty when identically zero
    std::vector<double> Fmagn, dFmagn_dh, d2Fmagn_dh2;  // NS [Nx]; empty for BH
};

// One tabulated passband prepared for the kernels (EnergyPassband, cpp/include/passband.hpp:26-47): per sample
//   a_j = (c h / k_B) / lambda_j                    the exponent of B_lambda is a_j / T
//   w_j = dlambda_j * transmission_j * 2 h c^2 / lambda_j^5,  dlambda_j = the trapezoid weight of util.cpp:21-30
//         (lambda_1 - lambda_0, lambda_{j+1} - lambda_{j-1}, lambda_{n-1} - lambda_{n-2}; the two factors 1/2 of the
//          nested trapezoids are applied once at the end)
// and t_dnu = trapz(lambdas, transmission * c / lambda^2) (cpp/src/passband.cpp:64-68, passband.hpp:41-42).
struct PassbandTable {
    std::vector<double> a, w;
    double t_dnu = 0.;
};
int prepare_passband(const double* lambdas, const double* transmissions, int n, PassbandTable& out, std::string& err);

// Returns FB200_OK or the error class the reference's constructor would have thrown, message in err.
int resolve(const fb200_args_t& a, Resolved& out, std::string& err);

// What an ensemble has to know about a model before it is resolved (allocation plan of the construction pipeline): read off the
// argument struct alone, and checked against the resolved model when that arrives.
struct ModelPlan {
    int Nx, Nt;
    bool is_ns, has_A, has_B, has_C, dynamic;  // static wind arrays A / B / C (+ d2Fmagn/dh2 for NS), dynamic wind
};
ModelPlan plan_of(const fb200_args_t& a);
// the semi-axis resolve() stores in fb200_model_dev_t::star_semiaxis (key of the companion-star geometry tables)
double star_semiaxis_of(const fb200_args_t& a);

}  // namespace fb200
