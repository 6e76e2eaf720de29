// Companion-star geometry for the Fnu*_star columns (--starflux): the triangulated Roche-lobe surface the
// reference builds once per model in the FreddiState constructor.  Host-side set-up only; the per-row irradiation
// and the spectral sums run in the kernels (disc_engine.h, star_row).
//
// Stands in for (paths relative to the reference repository):
//   unit_sphere_triangles, Triangle            cpp/src/geometry.cpp:126-251   (icosahedron cpp/include/geometry.hpp:104-121)
//   RochePotential / CriticalRocheLobe / ...   cpp/src/rochelobe.cpp:19-134
//   Star::initializeRocheTriangles             cpp/src/star.cpp:29-40
//   FreddiState ctor                           cpp/src/freddi_state.cpp:74-81  (semiaxis: cpp/include/orbit.hpp:24-29)
#pragma once
#include <string>
#include <vector>

namespace fb200 {

// Per triangle: centre (3), unit normal (3), area — the only members Star::integrate and IrradiatedStar::Qirr read
// (cpp/src/star.cpp:50-64,246-258) — plus the vertices for the star dump of --fulldata.
struct StarGeometry {
    int n_tri = 0;
    std::vector<double> cx, cy, cz, nx, ny, nz, area;
    std::vector<double> vertices;  // [n_tri][3 vertices][3 coordinates]
    // packed for the device: cx | cy | cz | nx | ny | nz | area, n_tri doubles each
    std::vector<double> packed() const;
};

// semiaxis [cm], mass_ratio = Mopt / Mx, fill = rochelobefill, lod = starlod (20 * 4^lod triangles; lod <= 6)
int build_star_geometry(double semiaxis, double mass_ratio, double fill, int lod, StarGeometry& out, std::string& err);

}  // namespace fb200
