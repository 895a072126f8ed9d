/* amdg.h -- C ABI of the B200-native fast sparse-grid transform path of AdaM-DG.
 *
 * The reference (JuntaoHuang/adaptive-multiresolution-DG) has no FFI layer: its boundary is the C++ class
 * surface FastMultiplyLU / FastInterpolation / FastInitial / FastRHS (include/FastMultiplyLU.h:9-704),
 * LagrInterpolation / HermInterpolation (include/Interpolation.h:36-392) and ExplicitRK
 * (include/ODESolver.h:76-248).  Every entry point below names the reference function it replaces; the C++
 * mirror of those classes that calls this ABI lives in adaptive-multiresolution-dg_b200/host/amdg_host.hpp and
 * the reference-side binding is shown in INTEGRATION.md.
 *
 * Conventions (identical to the reference):
 *   - an element is (level[d], suppt[d]); element rows are in the caller's order (the glue passes them in
 *     DGSolution::dg iteration order, include/DGSolution.h:219);
 *   - a coefficient array is `double[n_comp][n_elem][edge^dim]`, each element block row-major with the last
 *     dimension fastest (include/VecMultiD.h:114-152); during a chain, swept dims have the target edge;
 *   - a 1D operator is dense row-major `mat[from_basis][to_basis]`, basis index = ord1d*(pmax+1)+p
 *     (source/FastMultiplyLU.cpp:479,497-499);
 *   - every call returns 0 on success or a negative AMDG_E* code; nothing exits the process
 *     (the reference prints and calls exit(1), e.g. source/FastMultiplyLU.cpp:268).
 * All `double*` / `const double*` arguments named dev_* are DEVICE pointers; host_* are host pointers.
 * There is no CPU fallback: compute calls on a context created without a device fail with AMDG_ENODEVICE.
 */
#ifndef AMDG_H
#define AMDG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct amdg_ctx amdg_ctx;

enum { AMDG_OK = 0, AMDG_EINVAL = -1, AMDG_ENODEVICE = -2, AMDG_ECUDA = -3, AMDG_ENOMEM = -4, AMDG_ESTATE = -5 };

/* relation of a sweep: Element::ptr_vol_alpt / ptr_flx_alpt (include/Element.h:152-155) */
enum { AMDG_REL_VOL = 0, AMDG_REL_FLX = 1 };
/* part of the 1D operator: "L" = strictly finer source, "U" = same-or-coarser source, "full"
 * (source/FastMultiplyLU.cpp:490-491) */
enum { AMDG_LU_L = 0, AMDG_LU_U = 1, AMDG_LU_FULL = 2 };
/* schedule of amdg_apply_tensor: the reference's literal 2^(d-1) chains (source/FastMultiplyLU.cpp:614-664) or the
 * same sum with shared prefixes/suffixes (fewer sweeps, identical up to summation order) */
enum { AMDG_SCHED_LITERAL = 0, AMDG_SCHED_SHARED = 1 };
/* point-wise flux kinds: FluxFunction namespace (include/Interpolation.h:396-437) and the Vlasov products
 * (source/Interplation.cpp:4451-4497, 4523-4573) */
enum { AMDG_FLUX_LINEAR = 0, AMDG_FLUX_BURGERS = 1, AMDG_FLUX_SIN = 2, AMDG_FLUX_COS = 3,
       AMDG_FLUX_BUCKLEY_X = 4, AMDG_FLUX_BUCKLEY_Y = 5, AMDG_FLUX_VLASOV_SMOOTH_E = 6 };
/* explicit Runge-Kutta schemes: ForwardEuler, RK2SSP, RK2Midpoint, RK3SSP (source/ODESolver.cpp:203-301) */
enum { AMDG_RK_EULER = 0, AMDG_RK_RK2SSP = 1, AMDG_RK_RK2MID = 2, AMDG_RK_RK3SSP = 3, AMDG_RK_RK3HEUN = 4 };

const char *amdg_version(void);
const char *amdg_last_error(void);

/* ---- context: snapshot of the reference's statics Element::DIM, PMAX_alpt, PMAX_intp (include/Element.h:21-24),
 * DGSolution NMAX.  device = CUDA ordinal, or -1 for a host-only context (table building only). ---- */
int amdg_ctx_create(int dim, int nmax, int pmax_alpt, int pmax_intp, int device, amdg_ctx **out);
int amdg_ctx_destroy(amdg_ctx *ctx);
int amdg_ctx_set_stream(amdg_ctx *ctx, void *cuda_stream);   /* cudaStream_t; default: a stream owned by ctx */
int amdg_ctx_sync(amdg_ctx *ctx);
int amdg_ctx_info(amdg_ctx *ctx, int *out5);                 /* out5 = dim, nmax, pmax_alpt, pmax_intp, device */
int amdg_ctx_set_schedule(amdg_ctx *ctx, int sched);
int amdg_ctx_set_kernel(amdg_ctx *ctx, int variant);          /* 0 = auto: per launch the column kernel (one thread per column, FP64 FMA) for blocks of >= 64 doubles with KF*KT <= 9 and >= 32 columns, the lean tensor-core kernel otherwise, the whole-fibre tensor-core kernel for tiny sweeps,
                                                                 and -- adaptive mode -- the list-free gather kernel for tiny sweeps on a grid that replaced a short-lived one (DGAdapt::refine / coarsen every step);
                                                                 1 = gather, 2 = fibre-staged list kernel, 3 = pipelined list kernel, 4 = whole-fibre tensor-core (FP64 MMA) kernel, 5 = lean tensor-core kernel, 8 = column kernel
                                                                 (1-4 are independent implementations kept for the parity tests; 6 and 7 were measured 2.5-6x slower and removed, profiles/r02_sweep_kernels.md) */
int64_t amdg_ctx_launch_count(amdg_ctx *ctx);                 /* kernels launched so far by this context */
/* profiling aid: device buffer of n_items*8 int64 that the sweep kernel fills with per-CTA clock64 stamps (NULL = off) */
int amdg_ctx_set_debug_buffer(amdg_ctx *ctx, void *dev_buf);
/* diagnostic, host only (no device needed): build the work plans of the default sweep kernel for every fibre shape of dimension t
 * (a sweep with source block edges sizes_from[dim] and an operator kf -> kt) and verify their invariants; out[6] = shapes, pieces,
 * streamed (coarse) pieces, tile entries, largest staged row count, largest shared-memory need in doubles */
int amdg_lean_plan_check(amdg_ctx *ctx, int t, const int *sizes_from, int kf, int kt, int rel, int lu, int64_t *out);
/* ---- Hash (source/Hash.cpp:55-114) and 1D element order (source/Element.cpp:388-391), bit exact ---- */
int amdg_hash_key(int dim, const int *level, const int *suppt);
int amdg_order_elem(int level, int suppt);
/* Initial grid of DGSolution (source/DGSolution.cpp:10-57): fills level/suppt ([n][dim], construction order);
 * call with level == NULL to get the count. */
int64_t amdg_sparse_grid(int dim, int level_init, int sparse, int *level, int *suppt);
/* The grid of a field solution with auxiliary dimensions (DGSolution's second constructor, source/DGSolution.cpp:59-116): full grid of level_init in the
 * first dim - aux_dim dimensions, level 0 in the rest -- where E / B live next to f (example/07_vlasov_*.cpp).  Same calling convention. */
int64_t amdg_aux_grid(int dim, int level_init, int aux_dim, int *level, int *suppt);

/* ---- grid: replaces DGSolution::find_ptr_vol_alpt / find_ptr_flx_alpt (source/DGSolution.cpp:675-728) and
 * the per-adapt updates (source/DGAdapt.cpp:1073-1248).  Call after construction / refine / coarsen. ---- */
int amdg_grid_set(amdg_ctx *ctx, int64_t n_elem, const int *level, const int *suppt);
int64_t amdg_grid_size(amdg_ctx *ctx);
/* table export for parity: per element hash_key[n] and ord1d[n][dim] */
int amdg_grid_keys(amdg_ctx *ctx, int *hash_key, int *ord1d);
/* relation CSR along dim t: ptr[n+1], idx[ptr[n]] (element rows, ascending).  idx == NULL -> returns nnz. */
int64_t amdg_grid_relation(amdg_ctx *ctx, int t, int rel, int64_t *ptr, int *idx);
/* fibres along dim t: ptr[n_fibre+1], elems[n] sorted by ord1d inside a fibre.  elems == NULL -> returns n_fibre. */
int64_t amdg_grid_fibres(amdg_ctx *ctx, int t, int64_t *ptr, int *elems);

/* ---- operators: the dense OperatorMatrix1D tables (include/OperatorMatrix1D.h:22-121) and the transposed
 * point tables of FastLagrIntp / FastHermIntp (source/FastMultiplyLU.cpp:1316-1360) are compacted into their
 * (source ord1d, target ord1d) blocks.  edge_from/edge_to = pmax+1 of the row/column basis. ---- */
int amdg_op_register(amdg_ctx *ctx, const double *host_dense, int rows, int cols, int edge_from, int edge_to, int *op_out);
/* the canonical enumeration of related 1D element pairs (depends on nmax only): pairs grouped by target ord1d,
 * sources ascending; src/tgt/is_vol[n_pairs].  src == NULL -> returns n_pairs.  Blocks of a compact operator are
 * stored in this order. */
int64_t amdg_pairs(amdg_ctx *ctx, int *src_ord, int *tgt_ord, int *is_vol);
/* register an operator from its compact form blocks[n_pairs][edge_from][edge_to] (amdg_pairs order); the inverse,
 * amdg_op_blocks, exports the compact form of a registered operator. */
int amdg_op_register_compact(amdg_ctx *ctx, const double *host_blocks, int64_t n_pairs, int edge_from, int edge_to, int hier, int *op_out);
int amdg_op_blocks(amdg_ctx *ctx, int op, double *host_blocks);
/* hierarchisation stencils pwts (include/Interpolation.h:5-11, source/Interplation.cpp:775-887, 3166-3315):
 * anc[T-1][P1][2] = (ancestor ord1d, point index), wt[T-1][P1][P1] = wt[p0][ic], rows = 1D elements ord1d 1..T-1 */
int amdg_op_register_hier(amdg_ctx *ctx, const int *anc, const double *wt, int p1, int *op_out);
/* op = alpha*op_a + beta*op_b (e.g. ulft_vjp + urgt_vjp, source/FastMultiplyLU.cpp:1165) */
int amdg_op_combine(amdg_ctx *ctx, int op_a, double alpha, int op_b, double beta, int *op_out);

/* ---- the same tables generated by the library itself, straight into the compact form (no dense table, no reference run): one block per related 1D
 * element pair, O(2^nmax * nmax) blocks instead of the O(4^nmax) basis pairs the reference's constructors evaluate.
 * amdg_op_generate: table `table` of OperatorMatrix1D<U, AlptBasis>(.., "period") (include/OperatorMatrix1D.h:124-264 over Basis::product_volume /
 *   product_edge_dis_v / product_edge_dis_u, source/Basis.cpp:62-235); U = basis_u of degree pmax_u (mesh case msh_case_u of LagrBasis::set_interp_msh01,
 *   source/LagrBasis.cpp:33-154; ignored for Alpert / Hermite), V = Alpert of the context's degree.  The UX_* / *_VXAVE / UX_V tables exist for U = Alpert only.
 * amdg_op_generate_points: the transposed point table of FastLagrIntp / FastHermIntp (source/FastMultiplyLU.cpp:1316-1360): Alpert coefficients -> values
 *   (derivative 0: Lag_pt_Alpt_1D / Her_pt_Alpt_1D, source/Interplation.cpp:16-99, 1886-1965) or first derivatives (1: Lag_pt_Alpt_1D_d1) at the points.
 * amdg_op_generate_hier: the stencils of set_pts_wts_1d_ada_Lag / _Her (source/Interplation.cpp:775-887, 3166-3315) as the operator I + W.
 * amdg_points_generate: the 1D point coordinate table (LagrBasis::intep_pt / HermBasis::intep_pt); copied to host_pts1d[2^nmax * (pmax+1)] if not NULL and
 *   installed like amdg_points_set on a device context whose pmax_intp equals pmax. ---- */
enum { AMDG_BASIS_ALPERT = 0, AMDG_BASIS_LAGRANGE = 1, AMDG_BASIS_HERMITE = 2 };
enum { AMDG_TAB_U_V = 0, AMDG_TAB_U_VX = 1, AMDG_TAB_ULFT_VJP = 2, AMDG_TAB_URGT_VJP = 3, AMDG_TAB_UJP_VJP = 4, AMDG_TAB_UAVE_VJP = 5, AMDG_TAB_UJP_VXLFT = 6,
       AMDG_TAB_UJP_VXRGT = 7, AMDG_TAB_UX_VX = 8, AMDG_TAB_UXAVE_VJP = 9, AMDG_TAB_UJP_VXAVE = 10, AMDG_TAB_UX_V = 11 };
int amdg_op_generate(amdg_ctx *ctx, int basis_u, int pmax_u, int msh_case_u, int table, int *op_out);
/* the same for the other boundary types of Basis::product_edge_dis_v / _u (source/Basis.cpp:127-235): "zero" (all discontinuity points, no wrap) and
 * "inside" (points on x = 0, 1 skipped); the volume tables do not depend on it */
enum { AMDG_BC_PERIOD = 0, AMDG_BC_ZERO = 1, AMDG_BC_INSIDE = 2 };
int amdg_op_generate_bc(amdg_ctx *ctx, int basis_u, int pmax_u, int msh_case_u, int table, int boundary, int *op_out);
int amdg_op_generate_points(amdg_ctx *ctx, int basis, int pmax, int msh_case, int derivative, int *op_out);
int amdg_op_generate_hier(amdg_ctx *ctx, int basis, int pmax, int msh_case, int *op_out);
int amdg_points_generate(amdg_ctx *ctx, int basis, int pmax, int msh_case, double *host_pts1d);

/* ---- K1: one 1D sweep, FastMultiplyLU::transform_1D (source/FastMultiplyLU.cpp:436-512).
 * sizes_from[dim] = block edges of src; dst has the same edges except edge_to(op) in dim t.
 * dst = coef * sweep(src) (+ dst if accumulate).  n_comp components, each n_elem*block apart. ---- */
int amdg_sweep1d(amdg_ctx *ctx, int op, int rel, int lu, int t, const int *sizes_from,
                 const double *dev_src, double *dev_dst, int n_comp, double coef, int accumulate);

/* the same sweep for n_job (src, dst) pairs in one launch: the jobs share operator, relation, L/U part, dimension and the edges of
 * the dims after t; sizes_from[n_job][dim], dev_src/dev_dst[n_job], coef[n_job], accumulate[n_job].  This is how the sweeps of one
 * level of the shared-prefix schedule are issued (source/FastMultiplyLU.cpp:614-664 lists them one by one); the fibre-partitioned
 * multi-GPU path drives its local phases through it. */
int amdg_sweep1d_batch(amdg_ctx *ctx, int op, int rel, int lu, int t, const int *sizes_from, const double *const *dev_src,
                       double *const *dev_dst, const double *coef, const int *accumulate, int n_job, int n_comp);

/* the same with mapped destinations (lean tensor-core and column kernels, one component per job): dev_dst_map[j] (or NULL) = per element row the
 * offset in doubles, relative to dev_dst[j], of the element's destination block -- possibly in peer memory: the last sweep before a layout
 * switch of the fibre-partitioned multi-GPU path stores every block straight into the rank that owns it next; dev_acc_from[j] (or NULL, needs
 * accumulate[j]) = array in the plain row layout whose values are added instead of the destination's ("remote = local partial sum + sweep") */
int amdg_sweep1d_batch_mapped(amdg_ctx *ctx, int op, int rel, int lu, int t, const int *sizes_from, const double *const *dev_src,
                              double *const *dev_dst, const double *coef, const int *accumulate, const int64_t *const *dev_dst_map,
                              const double *const *dev_acc_from, int n_job);

/* the same with a SECOND destination per job (or NULL): every output block is also stored at dev_dst2[j] + dev_dst2_map[j][row] (peer memory) -- a buffer that
 * is consumed in this layout and, after the next layout switch, in the other one leaves the producing sweep for both places (the column kernel stores twice
 * from its epilogue; after the other kernels the library copies the rows) */
int amdg_sweep1d_batch_dual(amdg_ctx *ctx, int op, int rel, int lu, int t, const int *sizes_from, const double *const *dev_src,
                            double *const *dev_dst, const double *coef, const int *accumulate, const int64_t *const *dev_dst_map,
                            const double *const *dev_acc_from, double *const *dev_dst2, const int64_t *const *dev_dst2_map, int n_job);

/* ---- sum over all orderings of the chain of sweeps: FastRHS::transform_fucoe_to_rhs (source/FastMultiplyLU.cpp:4-16),
 * FastInterpolation::transform_ucoealpt_to_upintp (:740-819), FastInitial::transform_ucoeintp_to_ucoealpt (:1418-1445).
 * ops[dim], rels[dim]; src blocks edge_from^dim, dst blocks edge_to^dim; dst = coef*(...) (+ dst if accumulate). ---- */
int amdg_apply_tensor(amdg_ctx *ctx, const int *ops, const int *rels, const double *dev_src, double *dev_dst,
                      int n_comp, double coef, int accumulate);

/* the *_coarse_grid forms of the same transforms: FastRHS::transform_fucoe_to_rhs_coarse_grid (source/FastMultiplyLU.cpp:18-30),
 * FastInterpolation::transform_ucoealpt_to_upintp_coarse_grid (:760-779) over FastMultiplyLU::transform_1D_coarse_grid (:514-594): elements whose
 * levels sum to more than mesh_nmax are skipped as sources and as targets in every sweep (their part of dst is 0, or untouched if accumulate). */
int amdg_apply_tensor_coarse(amdg_ctx *ctx, const int *ops, const int *rels, const double *dev_src, double *dev_dst,
                             int n_comp, double coef, int accumulate, int mesh_nmax);

/* ---- hierarchisation: LagrInterpolation::eval_fp_to_coe_D_Lag / eval_up_to_coe_D_Lag
 * (source/Interplation.cpp:1222-1430, 891-1048) and the Hermite twins (:3651-3849, 3319-3479); in place allowed ---- */
int amdg_hierarchize(amdg_ctx *ctx, int hier_op, const double *dev_src, double *dev_dst, int n_comp);

/* ---- K2: point-wise flux, LagrInterpolation::eval_fp_Lag (source/Interplation.cpp:256-295):
 * dev_fp[c] = f_{flux_id[c]}(dev_up) for c < n_flux, params[c][4].  dev_pts (optional) = coordinates
 * [n_elem][edge^dim][dim] of the interpolation points for the Vlasov products. ---- */
int amdg_pointwise(amdg_ctx *ctx, int n_flux, const int *flux_id, const double *params, const double *dev_up,
                   double *dev_fp, const double *dev_pts);
/* Hermite form (DIM == 2, HermBasis::PMAX == 3, scalar): HermInterpolation::eval_fp_Her_2D (source/Interplation.cpp:2045-2288):
 * value slots f(u), first-derivative slots f'(u) u_x, mixed slot f''(u) u_x u_y + f'(u) u_xy.  Flux kinds LINEAR..COS. */
int amdg_pointwise_hermite2d(amdg_ctx *ctx, int n_flux, const int *flux_id, const double *params, const double *dev_up, double *dev_fp);
/* interpolation point coordinates of every element point from the 1D table pts1d[T*(pmax_intp+1)]
 * (LagrBasis::intep_pt, source/LagrBasis.cpp:19-28) */
int amdg_point_coords(amdg_ctx *ctx, const double *host_pts1d, double *dev_pts);

/* point-wise expressions: LagrInterpolation::eval_fp_Lag with all VEC_NUM unknowns (source/Interplation.cpp:256-295), eval_coe_u_Lag with a
 * coefficient of position (:648-698, wrappers :4159-4225) and the Vlasov bodies (:4451-4497, 4523-4573) with the field values that
 * DGSolution::copy_up_intp_to_f (source/DGSolution.cpp:1024-1065) broadcasts.  The reference passes std::function objects; here every output is a
 * stack program prog[n_prog][2] = (operation, argument), output c = ops [out_ptr[c], out_ptr[c+1]):
 *   VAR v: point value dev_up[v][point];  X t: coordinate t of the point (amdg_points_set);  OTHER j: dev_other[j][map[element]][local point]
 *   (dev_other_map = element row of the field grid for every element, NULL = same rows);  CONST k: consts[k];  binary + - * / pow min max;
 *   unary neg sin cos sqr exp sqrt abs tanh.  Only outputs listed are written (the reference's is_intp mask). */
enum { AMDG_PW_VAR = 1, AMDG_PW_X = 2, AMDG_PW_OTHER = 3, AMDG_PW_CONST = 4, AMDG_PW_ADD = 5, AMDG_PW_SUB = 6, AMDG_PW_MUL = 7, AMDG_PW_DIV = 8,
       AMDG_PW_NEG = 9, AMDG_PW_SIN = 10, AMDG_PW_COS = 11, AMDG_PW_SQR = 12, AMDG_PW_EXP = 13, AMDG_PW_SQRT = 14, AMDG_PW_ABS = 15, AMDG_PW_POW = 16,
       AMDG_PW_TANH = 17, AMDG_PW_MIN = 18, AMDG_PW_MAX = 19 };
/* the 1D table of interpolation point coordinates pts1d[T*(pmax_intp+1)] (LagrBasis::intep_pt, source/LagrBasis.cpp:19-28), kept on the device */
int amdg_points_set(amdg_ctx *ctx, const double *host_pts1d);
int amdg_pointwise_expr(amdg_ctx *ctx, int n_var, const double *const *dev_up, int n_other, const double *const *dev_other, const int *dev_other_map,
                        int n_out, double *const *dev_out, const int *prog, int n_prog, const int *out_ptr, const double *consts, int n_const);

/* ---- K4: explicit RK stage, ExplicitRK::step_stage (source/ODESolver.cpp:209-301): updates dev_u in place ---- */
int amdg_rk_stage(amdg_ctx *ctx, int scheme, int stage, double dt, const double *dev_u_tn, double *dev_u,
                  const double *dev_rhs, int64_t n);
/* RK4ODE2nd::step_stage (source/ODESolver.cpp:578-615) for u_tt = L u as the pair (u, v = u_t): call with dev_rhs = L u for
 * stage 0..3; dev_ku / dev_kv are caller-owned scratch of 4*n doubles each (the k1..k4 of the reference). */
int amdg_rk4_ode2nd_stage(amdg_ctx *ctx, int stage, double dt, const double *dev_u_tn, const double *dev_v_tn, double *dev_u,
                          double *dev_v, const double *dev_rhs, double *dev_ku, double *dev_kv, int64_t n);
/* y = alpha*x + beta*y */
int amdg_axpby(amdg_ctx *ctx, int64_t n, double alpha, const double *dev_x, double beta, double *dev_y);
/* y = sum_{i<k} coefs[i]*x_i + beta*y (k <= 16): DGSolution rhs accumulated from several FastRHS calls (source/FastMultiplyLU.cpp:69-93) in one pass */
/* Vlasov coupling: velocity moments of f accumulated into the right-hand side of a field solution that lives on the elements with level 0 in the
 * n_vdim trailing (velocity) dimensions: DGAdapt::compute_moment_1D2V / _2D2V (source/DGAdapt.cpp:243-338), any number of leading dimensions.
 * dev_map[n_field] = element row in f of every field element (-1: no partner, skipped, as the reference's `iter_f != f.dg.end()`);
 * order[n_vdim] in {0, 1} (the reference allows one first-order factor; the product form used here also gives the mixed moment);
 * dev_f [n_elem of f][a^dim], dev_rhs_field [n_field][a^dim] (only the entries with velocity degree 0 are touched: rhs += weight * moment).
 * The companion broadcast DGSolution::copy_up_intp_to_f (source/DGSolution.cpp:1024-1065) is the dev_other_map of amdg_pointwise_expr. */
int amdg_moment(amdg_ctx *ctx, int64_t n_field, const int *dev_map, int n_vdim, const int *order, double weight, const double *dev_f, double *dev_rhs_field);
/* refinement / coarsening indicator of DGAdapt::indicator_norm (source/DGAdapt.cpp:1018-1030): dev_norm[e] = sum_{v < n_var} || dev_u[v][e][.] ||_2 over
 * the Alpert coefficient blocks (a^dim doubles) of the indicator variables (for "wave" problems the caller lists ucoe_ut as further variables);
 * the adapt decision then needs n_elem doubles from the device instead of every coefficient */
int amdg_indicator_norm(amdg_ctx *ctx, int n_var, const double *const *dev_u, double *dev_norm);
int amdg_lincomb(amdg_ctx *ctx, int64_t n, int k, const double *coefs, const double *const *dev_x, double beta, double *dev_y);

/* ---- host-buffer entry points (what the reference-facing classes call; H2D/D2H inside) ---- */
int amdg_host_apply_tensor(amdg_ctx *ctx, const int *ops, const int *rels, const double *host_src, double *host_dst,
                           int n_comp, double coef, int accumulate);
int amdg_host_sweep1d(amdg_ctx *ctx, int op, int rel, int lu, int t, const int *sizes_from,
                      const double *host_src, double *host_dst, int n_comp, double coef, int accumulate);
int amdg_host_hierarchize(amdg_ctx *ctx, int hier_op, const double *host_src, double *host_dst, int n_comp);
/* cfg2 round trip: Alpert -> point values -> hierarchical interpolation coefficients -> Alpert
 * (FastLagrIntp::eval_up_Lagr, eval_up_to_coe_D_Lag, FastLagrInit::eval_ucoe_Alpt_Lagr) */
int amdg_host_roundtrip(amdg_ctx *ctx, int op_alpt_to_pt, int hier_op, int op_intp_to_alpt,
                        const double *host_ucoe_in, double *host_ucoe_out, int n_comp);

/* ---- multi-GPU plumbing (one process per GPU; the fibre-partitioned path of SURVEY.md 8e): device memory of the other ranks mapped through CUDA IPC,
 * a copy of element rows to mapped destinations, and a device-side barrier over peer-mapped flags (graph-capturable; gives up and raises *dev_error
 * instead of hanging when a peer never arrives).  flag_ptrs[r] = rank r's flag array unsigned[world] as mapped on this device. ---- */
int amdg_peer_export(amdg_ctx *ctx, const void *dev_ptr, void *handle64);
int amdg_peer_open(amdg_ctx *ctx, const void *handle64, void **dev_ptr_out);
int amdg_peer_close(amdg_ctx *ctx, void *dev_ptr);
int amdg_peer_barrier(amdg_ctx *ctx, void *const *flag_ptrs, int world, int rank, void *dev_epoch, void *dev_error);
int amdg_scatter_rows(amdg_ctx *ctx, const double *dev_src, int64_t n_rows, int width, double *dev_dst_base, const int64_t *dev_map);

/* ---- device memory helpers (so callers without a CUDA runtime binding can stage data) ---- */
int amdg_dev_alloc(amdg_ctx *ctx, int64_t n_doubles, double **dev_out);
int amdg_dev_free(amdg_ctx *ctx, double *dev);
int amdg_dev_upload(amdg_ctx *ctx, double *dev_dst, const double *host_src, int64_t n_doubles);
int amdg_dev_download(amdg_ctx *ctx, double *host_dst, const double *dev_src, int64_t n_doubles);
int amdg_dev_zero(amdg_ctx *ctx, double *dev, int64_t n_doubles);

#ifdef __cplusplus
}
#endif
#endif /* AMDG_H */
