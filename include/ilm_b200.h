/* ilm_b200.h -- C ABI of libilm_b200.so, the B200 (sm_100a) implementation of
 * the ImmersedLayers.jl immersed-layer operator hot path.
 *
 * The reference (JuliaIBPM/ImmersedLayers.jl v0.5.6) has no FFI: its seam is
 * Julia dispatch on `BasicILMCache` (src/cache.jl:16-88).  Each entry point
 * below names the reference method whose body a Julia `ccall` shim replaces
 * (see INTEGRATION.md).  Conventions:
 *   - all arrays are fp64, column-major with x (first index) fastest, exactly
 *     the memory layout of CartesianGrids' Nodes/Edges/ScalarData/VectorData;
 *     Edges = [vec(u); vec(v)], VectorData = [u; v];
 *   - every data pointer may be a device pointer (zero-copy, asynchronous on
 *     the plan's stream) or a host pointer (staged H2D/D2H inside the call,
 *     the call returns after the result is in host memory);
 *   - a plan is bound to one CUDA stream and is not re-entrant (the
 *     reference's caches are shared mutable scratch, src/matrix_operators.jl:10-25);
 *   - functions return ILM_OK or an error code; ilm_last_error() gives text.
 *     The Julia shim maps ILM_ESIZE -> DimensionMismatch/AssertionError and
 *     ILM_EINVAL -> ArgumentError/MethodError (test/tools.jl:34-35,86).
 * There is no CPU fallback: without a CUDA device ilm_plan_create fails with
 * ILM_ECUDA.
 */
#ifndef ILM_B200_H
#define ILM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ilm_plan ilm_plan;

enum ilm_status { ILM_OK = 0, ILM_EINVAL = 1, ILM_ESIZE = 2, ILM_ECUDA = 3, ILM_ENCCL = 4, ILM_ENOMEM = 5 };

/* CartesianGrids DDF types listed at src/cache.jl:305.  Goza is named so that a caller gets a precise answer:
 * its kernel is defined only in the un-vendored CartesianGrids source (no formula, value or test of it exists in
 * ImmersedLayers.jl), so ilm_plan_create returns ILM_EINVAL ("Goza DDF not provided") instead of guessing. */
enum ilm_ddf { ILM_DDF_YANG3 = 0, ILM_DDF_M3 = 1, ILM_DDF_ROMA = 2, ILM_DDF_M4PRIME = 3, ILM_DDF_WITCHHAT = 4, ILM_DDF_GOZA = 5 };
/* GridScaling / IndexScaling (src/ImmersedLayers.jl:81, src/cache.jl:305-324) */
enum ilm_scaling { ILM_GRID_SCALING = 0, ILM_INDEX_SCALING = 1 };
/* staggered layouts */
enum ilm_layout {
    ILM_NODES_PRIMAL = 0, /* (NX-1) x (NY-1) at (i, j)         */
    ILM_NODES_DUAL = 1,   /* NX x NY         at (i-1/2, j-1/2) */
    ILM_XEDGES = 2,       /* Edges{Primal}.u NX x (NY-1) at (i-1/2, j) */
    ILM_YEDGES = 3,       /* Edges{Primal}.v (NX-1) x NY at (i, j-1/2) */
    ILM_EDGES = 4         /* [u; v] */
};
/* pointwise normal products fused into regularize / interpolate
 * (src/surface_operators.jl:98-343) */
enum ilm_normal_mode {
    ILM_NORMAL = 0, /* product!(V,n,f) / pointwise_dot!(f,n,V)                  */
    ILM_CROSS = 1   /* pointwise_cross!: V.u = n.v f, V.v = -n.u f / n.u V.v - n.v V.u */
};
/* Schur-complement builders of src/matrix_operators.jl */
enum ilm_schur {
    ILM_RTLINVR = 0,     /* create_RTLinvR      :9-30    */
    ILM_CLINVCT = 1,     /* create_CLinvCT      :40-61   */
    ILM_GLINVD = 2,      /* create_GLinvD       :135-155 */
    ILM_GLINVD_CROSS = 3 /* create_GLinvD_cross :195-215 */
};

/* PhysicalGrid as data: size(g), cellsize(g), origin(g) */
typedef struct {
    int NX, NY;   /* dual cells incl. ghosts */
    double dx;
    int I0x, I0y; /* 1-based index of the primal node at the physical origin */
} ilm_grid;

/* ---- plan = BasicILMCache -------------------------------------------------
 * ilm_plan_create replaces SurfaceScalarCache(...) -> _surfacecache,
 * _get_regularization, _get_laplacian (src/cache.jl:164-182, 232-261, 305-324):
 * scaled point coordinates, DDF window tables for the primal-node, dual-node,
 * u-edge and v-edge layouts, their cell-bucketed gather lists, the LGF
 * multiplier Ghat and all scratch fields, resident in HBM.
 *   x,y,nx,ny,ds : N surface points, unit normals, arc weights (host pointers)
 *   lgf          : nlgf x nlgf table G(i,j), column-major, nlgf >= max(NX,NY)
 *                  (CartesianGrids' LGF_TABLE or lgf.lgf_table(); host pointer)
 *   c0           : far-field constant subtracted by `L\w`
 *   lap_factor   : L.factor (1/dx^2 for GridScaling, src/cache.jl:323-324)
 *   stream       : cudaStream_t or NULL (default stream)                     */
int ilm_plan_create(const ilm_grid* grid, int N, const double* x, const double* y, const double* nx,
                    const double* ny, const double* ds, int ddf, int scaling, const double* lgf, int nlgf,
                    double c0, double lap_factor, void* stream, ilm_plan** plan);
/* A second cache on the same grid that shares the Laplacian, as `SurfaceScalarCache(shape, g; L = L, ...)` does
 * in the region caches of src/forcing.jl:201-248 and `PointCollectionCache` (src/cache.jl:269-292: pass nx = ny = 0,
 * ds = 1, ILM_GRID_SCALING for its weight-1 Regularize).  The child plan aliases the parent's multipliers, twiddle
 * tables, spectrum buffers and grid scratch and its stream; it owns only its points and DDF tables.  The parent must
 * outlive the child, kernels added to the parent afterwards are not seen by the child, and calls on parent and
 * children are serialised by the caller like calls on one plan.                                                  */
int ilm_plan_create_shared(ilm_plan* parent, int N, const double* x, const double* y, const double* nx,
                           const double* ny, const double* ds, int ddf, int scaling, ilm_plan** plan);
/* update_system (src/system.jl:26-50): new body positions, Ghat is kept */
int ilm_plan_update_points(ilm_plan* plan, int N, const double* x, const double* y, const double* nx,
                           const double* ny, const double* ds);
void ilm_plan_destroy(ilm_plan* plan);
/* The two full-size spectrum buffers of the convolution (2 Lx x NY complex each: 537 MB at 4096^2, 8.6 GB at 16384^2) are
 * allocated on the first full-grid convolution; this frees them again (e.g. after plan creation on a rank that only
 * runs slab solves, which keep 1/nranks of them). */
int ilm_plan_release_spectrum(ilm_plan* plan);
int ilm_plan_sync(ilm_plan* plan);
const char* ilm_last_error(void);
int ilm_plan_npoints(const ilm_plan* plan);
/* number of elements of a layout on this plan's grid */
int64_t ilm_layout_size(const ilm_plan* plan, int layout);
/* kernels of this library launched on the plan so far (bench bookkeeping) */
int64_t ilm_plan_launch_count(const ilm_plan* plan);

/* DDF table inspection (bit-exact parity tests).  W*W entries per point in
 * [k][b][a] order (a = x offset fastest); idx = 0-based column-major linear
 * index or -1 if the entry lies outside the field. */
int ilm_get_table(ilm_plan* plan, int layout, int* W, int64_t* idx, double* wR, double* wE);

/* ---- regularize! / interpolate! (src/surface_operators.jl:13-88) ---------- */
int ilm_regularize(ilm_plan* plan, int layout, const double* f, double* grid);
int ilm_interpolate(ilm_plan* plan, int layout, const double* grid, double* f);
/* ---- forcing regions: _apply_forcing! (src/forcing.jl:456-515) --------------
 * area region (:456-465): dy .+= str .* mask, one fused sweep (32 B per point; mask = NULL: whole domain, 24 B);
 * line / point region (:467-494): dy .+= R str accumulated on the cells under the DDF windows only (the reference
 * zero-fills a scratch field, regularizes into it and adds the whole field).  layout: ilm_layout incl. ILM_EDGES
 * (str = VectorData [u; v]).  dy is updated in place.                                                          */
int ilm_forcing_area_add(ilm_plan* plan, int layout, const double* str, const double* mask, double* dy);
int ilm_forcing_line_add(ilm_plan* plan, int layout, const double* str, double* dy);
/* regularize_normal!, regularize_normal_cross! (:98-103, :150-162) : Edges <- ScalarData */
int ilm_regularize_normal(ilm_plan* plan, int mode, const double* f, double* edges);
/* normal_interpolate!, normal_cross_interpolate! (:228-233, :270-291) : ScalarData <- Edges */
int ilm_normal_interpolate(ilm_plan* plan, int mode, const double* edges, double* f);

/* ---- staggered stencils with scaling (src/grid_operators.jl:25-134) ------- */
int ilm_divergence(ilm_plan* plan, const double* edges, double* nodes_primal);
int ilm_grad(ilm_plan* plan, const double* nodes_primal, double* edges);
int ilm_curl_n2e(ilm_plan* plan, const double* nodes_dual, double* edges);
int ilm_curl_e2n(ilm_plan* plan, const double* edges, double* nodes_dual);
int ilm_laplacian(ilm_plan* plan, int layout, const double* in, double* out);

/* ---- inverse_laplacian! (src/grid_operators.jl:153-179) -------------------
 * in place: w <- (G * w - c0 sum(w)) / L.factor on the layout's own index box.
 * ILM_EDGES solves u and v together (one complex transform).                */
int ilm_inverse_laplacian(ilm_plan* plan, int layout, double* w);
/* two independent fields of the given layouts in one complex transform */
int ilm_inverse_laplacian_pair(ilm_plan* plan, int layout1, double* w1, int layout2, double* w2);
/* extra convolution kernels on the same engine: exp(L,a) -> plan_intfact
 * (src/timemarching.jl:93,104 via ConstrainedSystems).  table: n x n, kernel id
 * returned in *id (id 0 is the inverse Laplacian). */
int ilm_add_kernel(ilm_plan* plan, const double* table, int n, double c0, double factor, int* id);
int ilm_convolve(ilm_plan* plan, int kernel_id, int layout, double* w);

/* ---- composite surface-grid operators (src/surface_operators.jl:357-725) -- */
int ilm_surface_divergence(ilm_plan* plan, int mode, const double* f, double* nodes_primal);
int ilm_surface_grad(ilm_plan* plan, int mode, const double* nodes_primal, double* f);
int ilm_surface_curl_s2n(ilm_plan* plan, int mode, const double* f, double* nodes_dual);
int ilm_surface_curl_n2s(ilm_plan* plan, int mode, const double* nodes_dual, double* f);
/* _get_mask! (src/surface_operators.jl:865-876): mask = -L^-1 D_s 1 on primal nodes */
int ilm_mask(ilm_plan* plan, double* nodes_primal);

/* ---- Schur-complement builders (src/matrix_operators.jl) ------------------
 * Columns [col_begin, col_end) of the N x N matrix, written column-major with
 * leading dimension N into A (N x (col_end-col_begin)).  Column ranges are the
 * multi-GPU shard.                                                          */
int ilm_create_schur(ilm_plan* plan, int which, double scale, int col_begin, int col_end, double* A);
/* Same probing with another convolution kernel between the pre- and post-operator (ids from
 * ilm_add_kernel; 0 = L^-1): -scale * E exp(L a) R is the Schur complement of an IF-HERK stage
 * (`S_i = -B2 H_i B1^T`, src/timemarching.jl:86-107 via ConstrainedSystems). */
int ilm_create_schur_kernel(ilm_plan* plan, int which, int kernel_id, double scale, int col_begin, int col_end,
                            double* A);
/* Same matrix as ilm_create_schur(ILM_RTLINVR) from the direct-table identity
 * S[k,l] = -(scale/factor) sum_p sum_q E[k,p] (G(|p-q|) - c0) R[q,l]: no
 * transform, O(N^2 W^4) look-ups.  Cross-check of the column-solve path and an optional fast
 * builder; it is NOT what bench.py's grid-point*solves/s metric times.                      */
int ilm_create_RTLinvR_direct(ilm_plan* plan, double scale, int col_begin, int col_end, double* A);
/* same direct form for an arbitrary n x n kernel table T (column-major, zero beyond n):
 * A = -scale/factor * E (T - c0) R, e.g. S_i = -E exp(L a) R of an IF-HERK stage from plan_intfact's table */
int ilm_create_schur_direct_kernel(ilm_plan* plan, const double* table, int n, double c0, double factor, double scale,
                                   int col_begin, int col_end, double* A);
int ilm_create_nRTRn(ilm_plan* plan, double scale, double* A);   /* :225-244 */
int ilm_create_surface_filter(ilm_plan* plan, double* C);        /* :254-268 */

/* ---- surface-point solve (user level: test/literate/dirichlet.jl:99,124) --
 * LU with partial pivoting (LAPACK getrf/getrs semantics, ipiv 1-based),
 * dense mat-vec power s <- C^k s.  `stream` as in ilm_plan_create.           */
int ilm_dense_factor(int n, double* A, int* ipiv, void* stream);
int ilm_dense_solve(int n, const double* LU, const int* ipiv, int nrhs, double* B, void* stream);
int ilm_dense_matvec_pow(int n, const double* C, int k, double* s, void* stream);

/* ---- vector-data cache: SurfaceVectorCache (src/cache.jl:184-202) ----------
 * The same plan serves both cache kinds (it holds the tables of all four
 * layouts).  VectorData = [u; v] (2N); TensorData = [dudx; dudy; dvdx; dvdy]
 * (4N); EdgeGradient{Primal,Dual} = [dudx; dudy; dvdx; dvdy] with dudx, dvdy on
 * primal nodes and dudy, dvdx on dual nodes (layout ILM_EDGEGRAD).  Tensor
 * component (i,j) = (velocity component, derivative direction); see DESIGN.md
 * section 3 for the (unpinned) pointwise tensor conventions.                   */
#define ILM_EDGEGRAD 5
enum ilm_tensor_mode { ILM_TENSOR_NORMAL = 0, ILM_TENSOR_SYMM = 1 };
/* regularize_normal!(EdgeGradient<-VectorData), regularize_normal_symm! (src/surface_operators.jl:113-138) */
int ilm_regularize_normal_tensor(ilm_plan* plan, int mode, const double* v, double* edgegrad);
/* normal_interpolate!(VectorData<-EdgeGradient), normal_interpolate_symm! (:242-267) */
int ilm_normal_interpolate_tensor(ilm_plan* plan, int mode, const double* edgegrad, double* v);
enum ilm_vs_op {
    ILM_VS_CROSS = 0, /* Nodes{Dual}:   n x v  /  n x (psi e_z)   (:172-178, :303-310) */
    ILM_VS_DOT = 1    /* Nodes{Primal}: n . v  /  n phi           (:188-200, :320-332) */
};
int ilm_regularize_normal_vs(ilm_plan* plan, int op, const double* v, double* nodes);
int ilm_normal_interpolate_vs(ilm_plan* plan, int op, const double* nodes, double* v);
/* regularize_normal_dot!(Edges<-TensorData) (:209-215), normal_dot_interpolate!(TensorData<-Edges) (:335-343) */
int ilm_regularize_normal_dot_tensor(ilm_plan* plan, const double* tau, double* edges);
int ilm_normal_dot_interpolate_tensor(ilm_plan* plan, const double* edges, double* tau);
/* grad!(EdgeGradient<-Edges), divergence!(Edges<-EdgeGradient) (src/grid_operators.jl:61-71) */
int ilm_grad_tensor(ilm_plan* plan, const double* edges, double* edgegrad);
int ilm_divergence_tensor(ilm_plan* plan, const double* edgegrad, double* edges);
/* surface_divergence!(Edges<-VectorData) / _symm (:559-591), surface_grad!(VectorData<-Edges) / _symm (:632-664) */
int ilm_vsurface_divergence(ilm_plan* plan, int mode, const double* v, double* edges);
int ilm_vsurface_grad(ilm_plan* plan, int mode, const double* edges, double* v);
/* surface_curl!(Nodes{Dual}<-VectorData) (:388-398), surface_curl!(VectorData<-Nodes{Dual}) (:443-452) */
int ilm_vsurface_curl_s2n(ilm_plan* plan, const double* v, double* nodes_dual);
int ilm_vsurface_curl_n2s(ilm_plan* plan, const double* nodes_dual, double* v);
/* _get_mask! with Edges grid data (vector cache) */
int ilm_mask_edges(ilm_plan* plan, double* edges);
/* mask!(w, cache) / complementary_mask!(w, cache) (src/surface_operators.jl:788-823) with
 * _scalar_mask_product! / _vector_mask_product! (:880-923): w (in place) is multiplied by the mask,
 * averaged onto w's layout (grid_interpolate!) when that is not the cache's own grid-data layout.
 *   scalar cache: ILM_NODES_PRIMAL, ILM_NODES_DUAL, ILM_XEDGES, ILM_YEDGES, ILM_EDGES
 *   vector cache: ILM_EDGES, ILM_NODES_DUAL, ILM_NODES_PRIMAL, ILM_EDGEGRAD                      */
enum ilm_cache_kind { ILM_SCALAR_CACHE = 0, ILM_VECTOR_CACHE = 1 };
int ilm_mask_product(ilm_plan* plan, int cache_kind, int layout, int complementary, double* w);
/* Schur builders on VectorData: 2N x 2N, columns [col_begin, col_end) */
enum ilm_schur_vector {
    ILM_V_RTLINVR = 0,    /* create_RTLinvR      (src/matrix_operators.jl:9-30)    */
    ILM_V_CLINVCT = 1,    /* create_CLinvCT      (:40-61)                          */
    ILM_V_CL2INVCT = 2,   /* create_CL2invCT     (:102-125), two inverse Laplacians */
    ILM_V_GLINVD = 3,     /* create_GLinvD       (:135-155)                        */
    ILM_V_GLINVD_SYMM = 4 /* create_GLinvD_symm  (:165-185)                        */
};
int ilm_create_schur_vector(ilm_plan* plan, int which, double scale, int col_begin, int col_end, double* A);
int ilm_create_nRTRn_vector(ilm_plan* plan, double scale, double* A);

/* ---- Helmholtz decomposition on the vector cache (src/helmholtz.jl) -------------------------------
 * v = curl(psi) + grad(phi) + vp with  L psi = -(masked curl v + R n x [v]),  L phi = masked div v + R n . [v].
 * ilm_helmholtz_potentials = vectorpotential_from_masked_curlv! (:84-96) and scalarpotential_from_masked_divv!
 * (:186-201) together: the jump terms are accumulated on the cells under the DDF windows and the TWO inverse
 * Laplacians ride one complex transform.  psi: Nodes{Dual}, phi: Nodes{Primal}; either output may be NULL (then its
 * right-hand side is ignored); a NULL right-hand side or NULL dv is a zero field (vectorpotential_from_curlv!,
 * scalarpotential_from_divv!, :114-124, 216-219).
 * ilm_vecfield_from_potentials = vecfield_from_vectorpotential! + vecfield_from_scalarpotential! + the sum of
 * vecfield_helmholtz! (:130-134, 228-232, 285-307) as one sweep; psi, phi, vp may be NULL.
 * ilm_vecfield_helmholtz = vecfield_helmholtz! (:285-307) end to end (potentials stay in plan scratch).
 * ilm_helmholtz_jump_add: out = in + sign * R (n x dv) (op ILM_VS_CROSS, Nodes{Dual}) or in + sign * R (n . dv)
 * (op ILM_VS_DOT, Nodes{Primal}): masked_curlv_from_curlv_masked! / masked_divv_from_divv_masked! (sign -1, :149-160,
 * 240-252) and curlv_masked_from_masked_curlv! / divv_masked_from_masked_divv! (sign +1, :171-181, 263-274).       */
int ilm_helmholtz_potentials(ilm_plan* plan, const double* masked_curlv, const double* masked_divv, const double* dv,
                             double* psi, double* phi);
int ilm_vecfield_from_potentials(ilm_plan* plan, const double* psi, const double* phi, const double* vp_edges, double* v_edges);
int ilm_vecfield_helmholtz(ilm_plan* plan, const double* masked_curlv, const double* masked_divv, const double* dv,
                           const double* vp_edges, double* v_edges);
int ilm_helmholtz_jump_add(ilm_plan* plan, int op, int sign, const double* dv, const double* in, double* out);

/* kernels launched by the three ilm_dense_* entry points since load (bench bookkeeping) */
int64_t ilm_dense_launch_count(void);

/* ---- convective terms (src/grid_operators.jl:258-434), each one fused sweep --------------------
 * The reference composes grad! / grid_interpolate! / product! / transpose! over EdgeGradient-sized
 * temporaries (ConvectiveDerivativeCache, RotConvectiveDerivativeCache); no extra cache is needed here.
 * Results are divided by dx for GridScaling (_scale_derivative!), except w x v (:404-410).           */
/* convective_derivative!(udp::Nodes{Primal}, u::Edges, p::Nodes{Primal}, ...) (:258-264, 318-327) */
int ilm_convective_derivative_scalar(ilm_plan* plan, const double* vel_edges, const double* nodes_primal, double* out_nodes_primal);
/* convective_derivative!(vdw::Nodes{Dual}, v::Edges, w::Nodes{Dual}, ...) (:272-288, 329-341) */
int ilm_convective_derivative_dual(ilm_plan* plan, const double* vel_edges, const double* nodes_dual, double* out_nodes_dual);
/* convective_derivative!(vdu::Edges, v::Edges, u::Edges, ...) (:290-299, 361-375); u_edges == vel_edges gives
 * convective_derivative!(udu, u, ...) (:308-316, 343-359)                                            */
int ilm_convective_derivative_vector(ilm_plan* plan, const double* vel_edges, const double* u_edges, double* out_edges);
/* w_cross_v!(vw::Edges, w::Nodes{Dual}, v::Edges, ...) (:404-434) */
int ilm_w_cross_v(ilm_plan* plan, const double* w_nodes_dual, const double* vel_edges, double* out_edges);

/* ---- slab decomposition of one convolution over several GPUs ---------------------------------
 * Multi-GPU form of inverse_laplacian! (src/grid_operators.jl:153-179; the reference itself is
 * single-process) for ONE grid whose rows are spread over `nranks` GPUs, one process per GPU
 * (BASELINE config C5b).  Rank r owns the field rows [row0,row1) and, for the
 * column pass, the x-frequency tile columns [tc0,tc1).  A solve is
 *     ilm_slab_forward -> all-to-all #1 -> ilm_slab_columns -> all-to-all #2 -> ilm_slab_inverse
 * where the library computes and packs / unpacks the exchange buffers and the host language issues
 * the two variable-count all-to-alls (counts from ilm_slab_counts, in doubles, peer-major buffers).
 * ilm_slab_partition / _counts / _buffer_doubles are pure host arithmetic (no GPU needed).
 * Every data pointer of the three compute calls is a DEVICE pointer; w*_rows address this rank's first
 * row (row0) of the field, mx * (min(row1, my) - row0) doubles.                                    */
typedef struct {
    int nranks, rank;
    int Lx, Ly;     /* half padded transform lengths                                   */
    int MYp;        /* rows carried through the spectrum (even)                        */
    int row0, row1; /* field rows of this rank (even boundaries; row1 may exceed my by the pad row) */
    int ntc;        /* tile columns (pairs of x-frequency columns) in total            */
    int tc0, tc1;   /* tile columns of this rank                                       */
    int cpw;        /* x-frequency columns per work item of the column pass            */
    int wlo, whi;   /* work items of the column pass owned by this rank                */
} ilm_slab_info;
/* rows = number of rows (my) of the widest field carried by the solve */
int ilm_slab_partition(int NX, int NY, int rows, int nranks, int rank, ilm_slab_info* out);
/* phase 0: exchange #1 (rows -> columns), phase 1: exchange #2 (columns -> rows); send[q], recv[q] in doubles */
int ilm_slab_counts(int NX, int NY, const ilm_slab_info* me, int phase, int64_t* send, int64_t* recv);
/* capacity (doubles) that each of the send and receive buffers needs */
int64_t ilm_slab_buffer_doubles(const ilm_slab_info* me);
int ilm_slab_forward(ilm_plan* plan, const ilm_slab_info* me, int layout1, const double* w1_rows, int layout2,
                     const double* w2_rows /* NULL: one field */, double* sendbuf);
int ilm_slab_columns(ilm_plan* plan, const ilm_slab_info* me, int kernel_id, const double* recvbuf, double* sendbuf);
int ilm_slab_inverse(ilm_plan* plan, const ilm_slab_info* me, const double* recvbuf, int layout1, double* w1_rows, int layout2,
                     double* w2_rows /* NULL: one field */);

/* ---- multi-GPU: NCCL behind the ABI (the reference is a single process) -----------------
 * One process per GPU.  Rank 0 obtains a 128-byte unique id and the host language distributes it (MPI.jl,
 * torch.distributed store, a file ...); every rank then binds a communicator to its plan.  All collectives of the
 * sharded entry points run on the plan's stream.  NCCL is resolved at run time (dlopen of libnccl.so.2): a
 * missing or failing NCCL gives ILM_ENCCL, the single-GPU entry points do not need it.                        */
int ilm_comm_unique_id(void* id128, int nbytes);
int ilm_comm_init(ilm_plan* plan, const void* id128, int nbytes, int rank, int nranks);
int ilm_comm_destroy(ilm_plan* plan);
int ilm_comm_info(const ilm_plan* plan, int* rank, int* nranks);
/* columns [lo, hi) of an n-column Schur build owned by `rank`: contiguous, even-sized blocks (column pairs share
 * one complex transform); pure host arithmetic                                                              */
int ilm_column_range(int n, int nranks, int rank, int* lo, int* hi);
/* create_RTLinvR / create_CLinvCT / create_GLinvD[_cross] (src/matrix_operators.jl:9-215) with the column loop
 * (:16-26) sharded over the plan's communicator: each rank probes its block into its slice of the device matrix,
 * one grouped in-place broadcast completes it; A (N x N, host or device) is the full matrix on every rank.
 * Without a communicator: all columns locally.                                                              */
int ilm_create_schur_sharded(ilm_plan* plan, int which, int kernel_id, double scale, double* A);
/* The whole Dirichlet Poisson problem `solve(prob, sys)` of test/literate/dirichlet.jl:71-107 in one call:
 * f* = L^-1 D_s [f], S = -E L^-1 R (sharded when the plan has a communicator), S s~ = (f+ + f-)/2 - E f*,
 * s = -s~, f = L^-1 R s + f*.  fplus, fminus (NULL = 0): N surface values; f: Nodes{Primal}; s: N; S_out (NULL or
 * N x N): the Schur complement.  Host or device pointers; S, its LU factors and every intermediate field stay on
 * the device, so a host caller moves 2N doubles in and one field + N doubles out.  f may be NULL: the rank wants
 * the multiplier only (sharded solve with the field returned on one rank); the last regularize + L^-1 are skipped. */
int ilm_dirichlet_poisson(ilm_plan* plan, const double* fplus, const double* fminus, double* f, double* s, double* S_out);
/* The same with the field returned for the grid rows [row0, row1) only (f_rows: (row1 - row0) rows of Nodes{Primal}, x fastest;
 * host or device): in a sharded solve every rank keeps a slab of the result and copies 1/nranks of the field to its host.
 * The reference has no counterpart (single process); rows of the union equal ilm_dirichlet_poisson's field bit for bit.     */
int ilm_dirichlet_poisson_rows(ilm_plan* plan, const double* fplus, const double* fminus, double* f_rows, int row0, int row1,
                               double* s, double* S_out);
/* One inverse Laplacian of row-distributed fields over the plan's communicator: ilm_slab_forward -> exchange ->
 * ilm_slab_columns -> exchange -> ilm_slab_inverse with both exchanges issued inside the library (grouped
 * ncclSend / ncclRecv); w*_rows are this rank's rows [row0, row1) (device), sendbuf / recvbuf device scratch of
 * ilm_slab_buffer_doubles(me) doubles each.                                                                 */
int ilm_slab_solve(ilm_plan* plan, const ilm_slab_info* me, int kernel_id, int layout1, double* w1_rows, int layout2,
                   double* w2_rows /* NULL: one field */, double* sendbuf, double* recvbuf);

/* ---- measurement hook ------------------------------------------------------
 * Times each pass of the convolution (A: rows forward, B: columns, C: rows
 * inverse) on the plan's stream with CUDA events: `reps` back-to-back launches
 * per pass on a two-field (complex) problem of the given layout; ms[i] is the
 * average launch duration of pass i.  The spectrum buffers (hundreds of MB at
 * 4096^2) exceed the L2, so no flush is needed between launches.            */
int ilm_profile_conv(ilm_plan* plan, int layout, int reps, double ms[3]);
/* Same, in Schur-probe mode: the right-hand sides are R e_col and R e_{col+1} (W x W patches), i.e. the
 * launches ilm_create_schur(ILM_RTLINVR) issues: pass A transforms only the patch rows, pass B uses the
 * sparse forward half transform.                                                                     */
int ilm_profile_conv_probe(ilm_plan* plan, int col, int reps, double ms[3]);
/* Row range [row_begin, row_end) of the probed grid field that ilm_create_schur(which) carries through
 * the inverse transforms (the rows under the interpolation windows, widened by the stencil reach); the
 * other rows are neither stored by the column pass nor inverted by the last row pass.             */
int ilm_probe_output_rows(ilm_plan* plan, int which, int* row_begin, int* row_end);

#ifdef __cplusplus
}
#endif
#endif /* ILM_B200_H */
