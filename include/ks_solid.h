/*
 * ks_solid.h — C ABI of the B200-native gsElasticity solid assembly path.
 *
 * Replaces the bodies of the solid closures of the reference
 *   Jacobian_t:  assembler.assemble(x, fixedDofs); m = assembler.matrix();   tutorials/nonlinear_solid_static.cpp:101-106
 *   Residual_t:  assembler.assemble(x, fixedDofs); v = assembler.rhs();      tutorials/nonlinear_solid_static.cpp:109-114
 *   ALResidual:  Force - lam*Force - rhs()                                    benchmarks/benchmark_Elasticity_Beam_APALM.cpp:307-325
 * for gsElasticityAssembler<real_t> on ONE trivariate tensor-product B-spline patch (total Lagrangian,
 * MaterialLaw = saint_venant_kirchhoff / neo_hooke_ln / neo_hooke_quad; `hooke` = the linear assemble()).
 * Both closures of the reference run the full assemble (K is built twice per Newton iteration);
 * ks_assemble produces K and rhs in one pass and the single-output entry points skip the other half.
 *
 * Conventions (shared with kl_shell.h): real_t = double, index_t = int32_t; control points numbered
 * i = i1 + n1*(i2 + n2*i3); DoFs component-major, free first, eliminated after ALL free DoFs
 * (index - n_free addresses fixed_values); the matrix is compressed-column, layout-compatible with
 * gsSparseMatrix<real_t>; 0 on success, negative KL_E_* code (kl_shell.h) on failure; nothing throws.
 */
#ifndef KS_SOLID_H
#define KS_SOLID_H

#include <stdint.h>
#include "kl_shell.h"

#ifdef __cplusplus
extern "C" {
#endif

/* gsElasticity material_law::law (options().setInt("MaterialLaw", ...), tutorials/nonlinear_solid_static.cpp:96) */
enum { KS_LAW_HOOKE = 0, KS_LAW_SVK = 1, KS_LAW_NEO_HOOKE_LN = 2, KS_LAW_NEO_HOOKE_QUAD = 3 };
/* G+Smo boxSide in 3-D: west(u=0) east(u=1) south(v=0) north(v=1) front(w=0) back(w=1);
 * corner k: bit0 = east, bit1 = north, bit2 = back (southwestfront = 0 ... northeastback = 7) */
enum { KS_WEST = 0, KS_EAST = 1, KS_SOUTH = 2, KS_NORTH = 3, KS_FRONT = 4, KS_BACK = 5 };

typedef struct ks_bc {
    int32_t side[6][3];    /* 1 = homogeneous/inhomogeneous Dirichlet on [face][component] (condition_type::dirichlet) */
    int32_t corner[8][3];  /* 1 = addCornerValue(corner, value, 0, component)                                         */
} ks_bc;

/* what the reference passes to gsElasticityAssembler<real_t>(ori, basis, bc, body_force) + options
 * (tutorials/nonlinear_solid_static.cpp:92-97) */
typedef struct ks_problem {
    int32_t degree[3];
    int32_t n_knots[3];
    const double* knots[3];        /* open knot vectors                                                    */
    const double* cp;              /* [n_cp*3] undeformed control net, xyz interleaved                     */
    const double* weights;         /* must be NULL: polynomial B-spline geometry only in this version     */
    const int32_t* dof_map;        /* [3*n_cp]                                                             */
    int32_t n_free, n_fixed;
    const double* fixed_values;    /* [n_fixed] or NULL (= 0): the fixedDofs of assemble(x, fixedDofs)     */
    int32_t material_law;          /* KS_LAW_*                                                             */
    double E, nu;                  /* "YoungsModulus", "PoissonsRatio"                                     */
    double body_force[3];          /* constant body force (gsFunctionExpr body_force)                      */
    int32_t n_tractions;           /* dead Neumann loads: condition_type::neumann on a face                */
    const int32_t* traction_side;  /* [n_tractions] KS_WEST..KS_BACK                                       */
    const double* traction_val;    /* [n_tractions*3] constant traction vector per face                    */
} ks_problem;

typedef struct ks_ctx ks_ctx;

/* DoF numbering of the component-wise gsDofMappers of gsElasticityAssembler (free DoFs of component 0, 1, 2, then the
 * eliminated ones in the same order) */
int ks_build_dofmap(int32_t n1, int32_t n2, int32_t n3, const ks_bc* bc, int32_t* dof_map, int32_t* n_free, int32_t* n_fixed);
/* gsElasticityAssembler ctor: tables, F_ext (body force + tractions), symbolic pattern on the GPU. KL_E_NOGPU without a device.
 * Performance note (results are identical either way): the tri-cubic Jacobian kernel keeps its 1-D basis tables in the 64 KB constant
 * memory of the device when the mesh has at most 85 elements per direction; ONE context per device owns those tables (the first such
 * ks_create takes them, ks_destroy releases them) and any other context alive at the same time runs the shared-memory instantiation,
 * about 13 % slower.  KS_NO_CONST=1 in the environment disables the constant tables. */
int ks_create(const ks_problem* prob, int device, ks_ctx** out);
void ks_destroy(ks_ctx* ctx);
int ks_sizes(const ks_ctx* ctx, int32_t* n_dofs, int64_t* nnz, int64_t* n_elements, int64_t* n_qp);   /* numDofs() */
int ks_pattern_host(const ks_ctx* ctx, int32_t* outer, int32_t* inner);
/* assemble(x, fixedDofs): K(x) -> values_host (may be NULL), rhs = F_ext - F_int(x) -> r_host (may be NULL); x NULL = 0 */
int ks_assemble(ks_ctx* ctx, const double* x_host, double* values_host, double* r_host);
int ks_jacobian(ks_ctx* ctx, const double* x_host, double* values_host);                 /* Jacobian_t body   */
int ks_residual(ks_ctx* ctx, const double* x_host, double* r_host);                      /* Residual_t body   */
int ks_al_residual(ks_ctx* ctx, const double* x_host, double lam, double* r_host);       /* F_int - lam F_ext */
int ks_force(ks_ctx* ctx, double* f_host);                                               /* assemble(); rhs() */
/* gsMassAssembler<real_t>(ori, basis, bc, body_force) with options "Density": assemble(); matrix()
 * (tutorials/nonlinear_solid_dynamic.cpp:98-109) on the pattern of K: M_ab^{cd} = delta_cd density int N_a N_b */
int ks_mass(ks_ctx* ctx, double density, double* values_host);
/* device-resident variant: x_dev / r_dev device pointers (r_dev may be NULL), matrix stays at ks_values_device */
int ks_assemble_device(ks_ctx* ctx, const double* x_dev, int want_matrix, double* r_dev, void* stream);
double* ks_values_device(ks_ctx* ctx);
int ks_check(ks_ctx* ctx, void* stream);          /* maps the device error flag (det F <= 0, non-finite) to KL_E_* */
int ks_last_timing(const ks_ctx* ctx, float* ms_points, float* ms_jacobian, float* ms_residual);
int ks_kernel_launches(const ks_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
