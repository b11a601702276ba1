// TEST INFRASTRUCTURE ONLY (see oracle.h).  Block preconditioners: see precond.cc.
#pragma once
#include <array>
#include <memory>
#include <vector>

#include "../ampe_b200/csrc/mg_cell.h"
#include "oracle.h"

namespace oracle {

struct Ctx;

// restated reference operators (independent of the product's arithmetic)
void elliptic_apply(const Box& b, const double* dx, View m, View c, View* d, const double* u, double* out);
void quat_stencil_apply(const Box& b, const double* h, double gamma, View sqrt_m, View* fc, const double* w,
                        double* out);

// host loop over the product's per-cell multigrid functions (ampe_b200/csrc/mg_cell.h); same
// interface and cycle structure as the device solver behind ampe_mg_* (ampe_b200/csrc/mg.cu)
class HostMG
{
 public:
   // ncomp components solved together with one matrix (the quaternion block: qlen)
   HostMG(int ndim, const int* n, const double* dx, bool with_s, int ncomp = 1);
   int numComponents() const { return d_nc; }
   void setElliptic(const double* m, int ngm, double m_const, const double* c, int ngc, double c_const,
                    const double* const* d, const double* const* d2, int ngd, double d_scale, double d_const);
   void setQuat(double gamma, const double* mobility, int ngm, const double* const* face_coef, int ngfc);
   void solve(const double* rhs, double* soln, int ncycles, bool symmetrized);
   void apply(const double* u, double* out) const;
   void setSweeps(int pre, int post, int coarse) { d_pre = pre, d_post = post, d_coarse = coarse; }
   // homogeneous Neumann boundary per direction (same rule as ampe_mg_set_zero_slope); before the set* calls
   void setZeroSlope(const int* zero_slope);
   // one red-black sweep per pass over tiles (mg_rb_tile_pass), same tile choice as ampe_b200/csrc/mg.cu;
   // min_cells: levels with fewer cells keep the colour half-sweeps (the device uses the tail threshold 4096)
   void setFused(bool on, long long min_cells = 4096);
   int fusedLevels() const;
   int numLevels() const { return (int)d_levels.size(); }
   const int* levelExtents(int l) const { return d_levels.at(l).n; }
   const double* levelArray(int level, int which) const;

 private:
   void buildCoarse();
   void configure(bool var_c, double c_const, bool var_m, double m_const, bool var_d, double d_const);
   void smooth(int l, int sweeps);
   void vcycle();
   int d_ndim;
   int d_nc = 1;
   bool d_with_s;
   bool d_set = false;
   int d_n[3];
   double d_inv_h2[3];
   int d_pre = 1, d_post = 1, d_coarse = 8;
   int d_zero_slope[3] = {0, 0, 0};
   std::vector<ampe_mg_cell::Level> d_levels;
   std::vector<std::vector<double>> d_store;
   std::vector<std::array<std::vector<double>, 5>> d_coef;
   mutable std::vector<double> d_scratch;
   std::vector<bool> d_two_colour;
   std::vector<ampe_mg_cell::TileShape> d_tile;
   bool d_fused_restriction = false;  // residual + restriction in one pass (mg_restrict_residual_cell)
   std::vector<std::vector<double>> d_alt_u;
};

int precond_setup(Ctx* c, double gamma, int ncycles, bool has_dquatdphi);
int precond_dquatdphi(Ctx* c, const double* z_phase, double* out);
void quatdiffusionderiv(const Box& b, double misorientation_factor, View temperature, View var, int depth,
                        View* gradq, View* diff, double gradient_floor, char smooth_floor_type, char interp_type,
                        char avg_type);
void quatmobilityderiv(const Box& b, View phase, View dmobility, double scale_mobility, double min_mobility,
                       char func_type, double alt_scale_factor);
void compute_dquatdphi_face_coef(const Box& b, View* dprime, View phi, View* fc);
void multicomponent_multiply(const Box& b, View factor, View var, int vnc);
int precond_solve(Ctx* c, const ampe_rhs_fields* r, const ampe_rhs_fields* z);
int precond_apply(Ctx* c, int block, const double* u, double* out);
HostMG* precond_block(Ctx* c, int block);
void precond_destroy(Ctx* c);

}  // namespace oracle
