/* pmb.h -- C ABI of libpmb.so: the B200 (sm_100a) hot path of pyMOTO's compliance design iteration.
 *
 * Every entry point replaces one numpy/scipy call site of the reference (pyMOTO v2.0.1, paths relative to
 * /root/reference) and is what a ctypes binding in the reference would call:
 *
 *   pmb_csr_pattern      pymoto/modules/assembly.py:130-206   (argsort/unique pattern build -> closed form)
 *   pmb_assemble         pymoto/modules/assembly.py:255-275   (np.add.at scatter, bc rows/cols, bc diagonal)
 *   pmb_assemble_sens    pymoto/modules/assembly.py:298-315 -> pymoto/common/dyadcarrier.py:408-412 (einsum)
 *   pmb_rowstats         pymoto/solvers/solvers.py:88-96      (get_diagonal_indices) + iterative.py:38-39 (diagonal)
 *   pmb_spmv             scipy csr_matvec at pymoto/solvers/iterative.py:236-255,359,375,382, solvers.py:84,237
 *   pmb_elem_spmv        the same call sites on the finest level, evaluated from x_e and Ke (assembly.py:255-261 folded in)
 *   pmb_smooth0          pymoto/solvers/iterative.py:43,233-234   (u = w r/D)
 *   pmb_restrict         pymoto/solvers/iterative.py:244      (R^T r, csc_matvec)
 *   pmb_prolong_add      pymoto/solvers/iterative.py:250      (u += R u_c, csr_matvec)
 *   pmb_galerkin         pymoto/solvers/iterative.py:173      (R^T A R, csr_matmat x2)
 *   pmb_galerkin_direct  the same call site on level 0, evaluated from x_e and Ke (+ pmb_scatter_add for the bc term)
 *   pmb_densify / pmb_dense_invert / pmb_dense_gemv
 *                        pymoto/solvers/sparse.py:533-550     (splu + solve on the coarsest operator)
 *   pmb_dots / pmb_lincomb / pmb_cg_xr_update
 *                        pymoto/solvers/iterative.py:376-395  (CG dot products and vector updates)
 *   pmb_bc_split         pymoto/solvers/solvers.py:175-176    (Dirichlet dofs: u = f/diag, rhs zeroed)
 *   pmb_pad_gather / pmb_stencil_corr / pmb_pad_scatter
 *                        pymoto/modules/filter.py:182-220     (FilterConv: index-mapped padding, scipy.signal convolve/correlate)
 *   pmb_filter_apply / pmb_vec_div
 *                        pymoto/modules/filter.py:266-270     (csc_matvec of H, division by Hs)
 *   pmb_oc_candidate     pymoto/common/optimizers.py:425-435  (OC bisection candidate)
 *   pmb_mma_asymptotes / pmb_mma_setup
 *                        pymoto/common/mma.py:129-140,178-217 (asymptote offsets; low/upp/alfa/beta/P/Q, rhs sums)
 *   pmb_mma_residual / pmb_mma_newton_sums / pmb_mma_newton_dir / pmb_mma_linesearch
 *                        pymoto/common/mma.py:313-336,349-392,401-423,428-462 (n-sized parts of subsolv)
 *   pmb_mma_gcmma_rho / pmb_mma_gcmma_estimate
 *                        pymoto/common/mma.py:151,236-239 (GCMMA: initial rho sums; approximation values and dk)
 *   pmb_pack_f32         pymoto/common/domain.py:541-548,579-583 (Float32 VTI payload, 2 -> 3 component padding)
 *   pmb_vcycle / pmb_pcg_solve
 *                        pymoto/solvers/iterative.py:222-256,340-403 (one V-cycle; the whole PCG solve driven from C)
 *
 * Conventions
 *   - all functions return 0 on success, non-zero on error; pmb_last_error() gives the (thread-local) message.
 *     Nothing throws across this boundary and nothing calls exit().
 *   - all array arguments are DEVICE pointers owned by the caller (FP64 unless stated); the library never frees
 *     them.  `stream` is a cudaStream_t passed as void*.
 *   - matrices are passed as the CSR `data` array only: on a structured voxel grid the CSR pattern is a closed
 *     form of (nx, ny, nz, ndof) (27-/9-point block stencil, rows and columns in node-major dof order), so
 *     `indptr`/`indices` are never read by the solver kernels.  pmb_csr_pattern materialises them bit-exactly
 *     for export.  `data` must be 16-byte aligned and padded by 2 doubles.
 *   - slab decomposition (multi-GPU): a rank owns node planes [kz0, kz0+nzl).  Nodal vectors, the bc mask and
 *     matrix rows are addressed relative to the first OWNED plane; halo planes kz0-1 and kz0+nzl are read at
 *     negative / past-the-end offsets and must be valid memory.  One GPU: kz0 = 0, nzl = nz+1.
 */
#ifndef PMB_H
#define PMB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  int nx, ny, nz; /* elements per direction; nz = 0 means 2-D (quad4)            */
  int ndof;       /* dofs per node, 1..3                                          */
  int kz0;        /* first owned node plane                                       */
  int nzl;        /* number of owned node planes                                  */
} pmb_grid;

/* coefficient c * (*num) / (*den or sqrt(*den)); num/den may be NULL (-> 1) and live in device memory */
typedef struct {
  double c;
  const double* num;
  const double* den;
  int sqrt_den;
} pmb_coef;

enum { PMB_SPMV = 0, PMB_RESIDUAL = 1, PMB_JACOBI = 2 };

const char* pmb_last_error(void);
int pmb_version(void);

/* number of matrix entries / rows owned by this slab */
long long pmb_nnz(const pmb_grid* g);
long long pmb_nrows(const pmb_grid* g);

/* K0: CSR pattern.  index_bits = 32 or 64 selects the integer type of indptr (nrows+1) and indices (nnz). */
int pmb_csr_pattern(const pmb_grid* g, void* indptr, void* indices, int index_bits, void* stream);

/* K1: data = sum_e x_e Ke (sequential adds from 0.0 in ascending element number, no FMA: bit-exact with
 * np.add.at), rows/cols in bcmask (1 byte per dof, may be NULL) zeroed, bc diagonal = bcdiagval.
 * x points at element layer kz0 (layer kz0-1 is read as halo when kz0 > 0). Ke is (nn*ndof)^2 row-major and a HOST
 * pointer (it is passed to the kernel through the parameter constant bank).  diag / nnz_offdiag (either may be NULL): the
 * row statistics of pmb_rowstats, produced while the rows are still on chip (no second pass over the values in 3-D). */
int pmb_assemble(const pmb_grid* g, const double* Ke_host, const double* x, const unsigned char* bcmask,
                 double bcdiagval, double* data, double* diag, int* nnz_offdiag, void* stream);

/* K11: dx_e = sum_{a,b} u[dof(e,a)] Ke[a,b] v[dof(e,b)], u and v taken as 0 at masked dofs.
 * Element layers [kz0, min(kz0+nzl, nz)) are produced (all elements in 2-D). accumulate != 0 adds into dx. */
int pmb_assemble_sens(const pmb_grid* g, const double* Ke, const double* u, const double* v,
                      const unsigned char* bcmask, double* dx, int accumulate, void* stream);

/* diag[r] = A[r,r]; nnz_offdiag[r] = number of non-zero off-diagonal entries in row r (int32). Either may be NULL. */
int pmb_rowstats(const pmb_grid* g, const double* data, double* diag, int* nnz_offdiag, void* stream);

/* K2/K3: y = A x | y = b - A x | y = x + w (b - A x)/diag.   y must not alias x.
 * dot_out (may be NULL): 3 doubles, receives sum_r y_r x_r, sum_r x_r dotv_r and sum_r y_r dotv_r over the
 * owned rows (the last two only when dotv != NULL); ws is a workspace of at least pmb_spmv_ws_doubles(g)
 * doubles used for the deterministic two-stage reduction. */
int pmb_spmv(const pmb_grid* g, int mode, const double* data, const double* x, const double* b,
             const double* diag, double w, double* y, const double* dotv, double* dot_out, double* ws,
             void* stream);
long long pmb_spmv_ws_doubles(const pmb_grid* g);
/* workspace (doubles, zero-initialised by the caller once) for pmb_dots / pmb_cg_xr_update */
long long pmb_ws_doubles(void);

/* Matrix-free application of the FINEST-level operator K = P (sum_e s_e Ke) P + bcdiagval (I - P) from the element
 * scaling vector s (the x that pmb_assemble was given: points at element layer kz0, layer kz0-1 read as halo) instead
 * of the assembled values.  Same modes, epilogues and fused dot products as pmb_spmv.  The operator is described by a
 * caller-owned POD (no state is kept in the library):
 *   Ke_host     (nn*ndof)^2 row-major, HOST pointer (passed to the kernel through the parameter constant bank)
 *   s, bcmask   DEVICE pointers (bcmask: 1 byte per dof, may be NULL); bcdiagval: diagonal of the Dirichlet rows
 *   brickflags  optional DEVICE bytes from pmb_elem_brickflags (layouts 4 - 7 skip all Dirichlet-mask traffic in bricks whose
 *               flag is 0; NULL = every brick checks the mask)
 *   variant     kernel layout, 0 .. pmb_elem_num_variants()-1 (3-D, ndof 1 or 3; everything else runs layout 0):
 *               0 = one node per thread on a 32x4x2 brick, 1 / 2 = z-marching 32x8 / 32x4 columns with ring-buffered
 *               planes, 3 = FP64 tensor-core (DMMA) layout for ndof = 3 (y equal to rounding), 4 / 5 = persistent CTAs
 *               (3 / 2 per SM) whose bricks are staged by TMA bulk copies into a 2-stage ring, 6 = FP64 tensor-core layout
 *               marching along y with in-register accumulation (ndof = 3, y equal to rounding).  Layouts 0, 1, 2, 4, 5
 *               give bit-identical y.
 * Layouts 4 - 6 copy whole 16-byte granules: the granules holding the first / last element of x and s must be readable
 * (true for any cudaMalloc'ed array; vectors from DeviceCSR.new_vec() are plane-padded).
 * ws: pmb_elem_ws_doubles(g) doubles. */
typedef struct {
  const double* Ke_host;
  const double* s;
  const unsigned char* bcmask;
  double bcdiagval;
  const unsigned char* brickflags;
  int variant;
} pmb_elem_op;
int pmb_elem_spmv(const pmb_grid* g, int mode, const pmb_elem_op* op, const double* x, const double* b, const double* diag,
                  double w, double* y, const double* dotv, double* dot_out, double* ws, void* stream);
long long pmb_elem_ws_doubles(const pmb_grid* g);
int pmb_elem_num_variants(void);
/* flags[unit] = 1 iff the region of the slab that layout `variant` (0, 4, 5: a 32x4x2-node brick + 1-node apron; 6, 7: one
 * 8-row step of a 32x2 node-column strip) stages for that unit of work holds a masked dof; pmb_elem_brickflags_bytes(g,
 * variant) bytes.  Computed once per bc set, slab and layout. */
long long pmb_elem_brickflags_bytes(const pmb_grid* g, int variant);
int pmb_elem_brickflags(const pmb_grid* g, int variant, const unsigned char* bcmask, unsigned char* flags, void* stream);
/* pmb_elem_autotune times every layout (Jacobi mode, y is scratch) on the caller's operands, stores the launch times in
 * ms_out[pmb_elem_num_variants()] and the fastest admissible layout in *best (the caller writes it into its pmb_elem_op):
 * allow_rounding = 0 admits only the layouts bit-identical to layout 0, != 0 also the tensor-core layouts (3, 6).
 * flags_scratch: pmb_elem_autotune_flag_bytes(g) bytes (brick flags are layout-specific and recomputed per layout; may be
 * NULL when op->bcmask is NULL).  Not capturable into a CUDA graph. */
long long pmb_elem_autotune_flag_bytes(const pmb_grid* g);
int pmb_elem_autotune(const pmb_grid* g, const pmb_elem_op* op, const double* x, const double* b, const double* diag, double* y,
                      unsigned char* flags_scratch, int allow_rounding, double* ms_out, int* best, void* stream);

/* Symmetric half-stencil storage of a coarse-level operator (whole 3-D grids; replaces the full-stencil sweeps of
 * pymoto/solvers/iterative.py:236-255 on the levels below the finest): S[slot][d][c][node], slot 0 = diagonal block, slots
 * 1..13 = the 13 upper neighbours, pmb_sym_doubles(g) doubles = 14/27 of the stencil-CSR values.  pmb_sym_pack copies the
 * blocks out of stencil-CSR `data` and leaves in stats[0..1] (device, 2 x uint64) the bit patterns of the doubles
 * max |A_ij - A_ji^T| and max |A_ij|: the caller must not use S when the first is not negligible against the second.
 * pmb_sym_spmv: modes / epilogues / fused dot products of pmb_spmv (ws: pmb_spmv_ws_doubles). */
long long pmb_sym_doubles(const pmb_grid* g);
int pmb_sym_pack(const pmb_grid* g, const double* data, double* S, unsigned long long* stats, void* stream);
int pmb_sym_spmv(const pmb_grid* g, int mode, const double* S, const double* x, const double* b, const double* diag, double w,
                 double* y, const double* dotv, double* dot_out, double* ws, void* stream);

/* u = w * (r / diag) */
int pmb_smooth0(long long n, double w, const double* r, const double* diag, double* u, void* stream);

/* K4: rc = R^T rf.  gf = fine grid (its kz0/nzl describe the fine slab), gc = coarse grid (coarse slab). */
int pmb_restrict(const pmb_grid* gf, const pmb_grid* gc, const double* rf, double* rc, void* stream);
/* K5: uf += R uc */
int pmb_prolong_add(const pmb_grid* gf, const pmb_grid* gc, const double* uc, double* uf, void* stream);
/* K6: Ac = R^T A R in the coarse grid's own stencil-CSR layout; work holds pmb_galerkin_ws_doubles(gf) doubles */
int pmb_galerkin(const pmb_grid* gf, const pmb_grid* gc, const double* Af, double* Ac, double* work, void* stream);
long long pmb_galerkin_ws_doubles(const pmb_grid* gf);
/* the two passes separately (multi-GPU: the lower halo plane of `work`, 27*ndof^2 doubles per node stored BEFORE
 * the pointer, is exchanged between them): cols = column collapse B = A R of the owned fine rows,
 * rows = row collapse Ac = R^T B of the owned coarse rows */
int pmb_galerkin_cols(const pmb_grid* gf, const pmb_grid* gc, const double* Af, double* work, void* stream);
int pmb_galerkin_rows(const pmb_grid* gf, const pmb_grid* gc, const double* work, double* Ac, void* stream);

/* K6 (direct): the level-1 operator straight from the element scaling vector of the finest level, Ac = sum_E sum_p
 * s_child(E,p) G_id(E,p) -- the fine matrix is never read (derivation and host set-up: pymoto_b200/coarse.py; replaces the two
 * csr_matmat of pymoto/solvers/iterative.py:173 on level 0).  Gtab: ntab tables of 8 x 8 x ndof x ndof doubles ([a][b][d][c];
 * entries 0..7 = unmasked children, the rest = children touching Dirichlet dofs); cidx (may be NULL): int32 per GLOBAL
 * coarse element, -1 or the row of child_ids (8 uint16 table indices per such element).  s as for pmb_assemble, but TWO
 * halo layers below the fine slab are read.  The Dirichlet diagonal term bcdiagval R^T (I-P) R is added afterwards by
 * pmb_scatter_add(n, idx, val, data): data[idx[i]] += val[i], idx unique. */
int pmb_galerkin_direct(const pmb_grid* gf, const pmb_grid* gc, const double* Gtab, const int* cidx, const unsigned short* child_ids,
                        const double* s, double* Ac, void* stream);
int pmb_scatter_add(long long n, const long long* idx, const double* val, double* data, void* stream);

/* K7: coarsest level. dense is n*n row-major. pmb_dense_invert inverts in place (blocked Gauss-Jordan without
 * pivoting, valid for SPD); scratch holds pmb_dense_invert_ws_doubles(n) doubles; info (device int) is set non-zero
 * on a non-positive pivot. */
int pmb_densify(const pmb_grid* g, const double* data, double* dense, void* stream);
int pmb_dense_invert(int n, double* dense, double* scratch, int* info, void* stream);
long long pmb_dense_invert_ws_doubles(int n);
int pmb_dense_gemv(int n, const double* M, const double* x, double* y, void* stream);

/* K8: out[i] = sum a_i . b_i for i < k (k <= 4), deterministic (fixed partial order). */
int pmb_dots(long long n, int k, const double* a0, const double* b0, const double* a1, const double* b1,
             const double* a2, const double* b2, const double* a3, const double* b3, double* out, double* ws,
             void* stream);

/* K9: out = ca*a + cb*b (b may be NULL); coefficients may reference device scalars. out may alias a or b. */
int pmb_lincomb(long long n, double* out, pmb_coef ca, const double* a, pmb_coef cb, const double* b,
                void* stream);
/* fused CG update: alpha = (*pr)/(*pq); x += alpha p; if q != NULL: r -= alpha q and rr_out = r.r */
int pmb_cg_xr_update(long long n, double* x, double* r, const double* p, const double* q, const double* pr,
                     const double* pq, double* rr_out, double* ws, void* stream);

/* Dirichlet split: sol = mask ? rhs/diag : 0 ; rhs_loc = mask ? 0 : rhs (mask: 1 byte per dof) */
int pmb_bc_split(long long n, const unsigned char* mask, const double* rhs, const double* diag, double* sol,
                 double* rhs_loc, void* stream);
/* mask[r] = (diag[r] != 0 && nnz_offdiag[r] == 0): rows whose only non-zero is the diagonal */
int pmb_diag_mask(long long n, const double* diag, const int* nnz_offdiag, unsigned char* mask, void* stream);
/* out = mask ? 0 : in */
int pmb_mask_zero(long long n, const unsigned char* mask, const double* in, double* out, void* stream);

/* K10: density filter stencil. wtab is the (2d+1)^3 ((2d+1)^2 in 2-D) weight table, z-major, x fastest.
 * out[e] = (sum_j w_ij in[j]) / (hs ? hs[e] : 1) over the window clipped to the domain, accumulated in
 * ascending element number with separate multiply and add (bit-exact with scipy csc_matvec).
 * in == NULL means in = 1 (row sums Hs).  Element layers [ez0, ez0+nezl) are produced; `in`, `hs`, `out`
 * point at layer ez0 and `in` is read d layers beyond on each side where those exist in the domain. */
int pmb_filter_apply(const pmb_grid* g, int ez0, int nezl, int d, const double* wtab, const double* in,
                     const double* hs, double* out, void* stream);
/* FilterConv (pymoto/modules/filter.py:8-220): padded convolution filter.
 * pmb_pad_gather: xpad (pz,py,px; x fastest) from x (nz,ny,nx) through per-axis index maps (-1 = constant padding with the
 * value cv*[index]; the outermost constant axis wins, z over y over x).  pmb_stencil_corr: out[o] = sum_q w[q] in[o+q-off]
 * with `in` zero outside its extent.  pmb_pad_scatter: dx[s] = sum of dxpad over the padded positions mapping to s
 * (per-axis inverse lists, CSR form: ptr (n+1), lst). */
int pmb_pad_gather(int nx, int ny, int nz, int px, int py, int pz, const int* mapx, const int* mapy, const int* mapz,
                   const double* cvx, const double* cvy, const double* cvz, const double* x, double* xpad, void* stream);
int pmb_pad_scatter(int nx, int ny, int nz, int px, int py, int pz, const int* ptrx, const int* lstx, const int* ptry,
                    const int* lsty, const int* ptrz, const int* lstz, const double* dxpad, double* dx, void* stream);
int pmb_stencil_corr(int inx, int iny, int inz, const double* in, int ox, int oy, int oz, double* out, int kx, int ky, int kz,
                     const double* w, int offx, int offy, int offz, void* stream);
/* out = a / b */
int pmb_vec_div(long long n, const double* a, const double* b, double* out, void* stream);

/* Multi-GPU halo mailboxes: two independent n-double copies in one launch (dst may be PEER memory mapped through
 * symmetric memory: the stores then travel over NVLink).  Any (src, dst) pair with a NULL member is skipped. */
int pmb_halo_copy2(long long n, const double* src0, double* dst0, const double* src1, double* dst1, void* stream);

/* ---- exchange steps over peer memory, ONE launch each (the Python host's default on z-slabs).  All pointers named "peer"
 * are this process's mappings of another rank's symmetric-memory allocation (e.g. torch symmetric memory, cuMemMap of an
 * exported handle): stores to them travel over NVLink.  No communicator, no host involvement, no separate barrier launch,
 * capturable in a CUDA graph.  Transport: every double travels as a 16-byte unit {low word, exchange number, high word, exchange
 * number}; the receiver polls its own memory until both numbers match (no fence, no flag round trip: one launch + one NVLink
 * store flight per exchange).  Every rank must issue the same sequence of these calls on one stream.  `ctl` is ordinary device
 * memory of the calling rank; mailboxes, tables and ctl are zeroed once (then a host barrier) before the first call.
 * pmb_peer_halo_exchange: send `n` doubles from send_lo / send_hi to the lower / upper neighbour and receive what they send
 * into recv_lo / recv_hi.  Any of the four data pointers may be NULL (one-way exchanges); a header unit is still exchanged with
 * both neighbours, which is what makes two alternating mailbox slots enough.  Replaces pack kernel + device barrier + unpack
 * kernel.  ctl[2] != 0 (also reduce ctl[1]): a neighbour never arrived within ~60 s -- the number of the first such exchange;
 * results are invalid from then on and later calls no longer wait. */
typedef struct pmb_peer_halo {
  double* box;     /* this rank's mailbox: pmb_peer_halo_box_doubles(cap) doubles, 16-byte aligned */
  double* box_lo;  /* the lower / upper neighbour's mailbox (peer pointers), NULL at the ends of the domain */
  double* box_hi;
  unsigned long long* ctl; /* 3 words of this rank's device memory: exchange counter, CTA counter, first timed-out exchange */
  long long cap;   /* largest n */
} pmb_peer_halo;
long long pmb_peer_halo_box_doubles(long long cap);
int pmb_peer_halo_exchange(const pmb_peer_halo* h, long long n, const double* send_lo, const double* send_hi,
                           double* recv_lo, double* recv_hi, void* stream);
/* pmb_peer_allreduce: in-place sum (op 0) or maximum (op 1) of count <= PMB_PEER_COUNT_MAX device doubles over all ranks: every
 * rank stores its values into its column of every rank's table, the columns are combined in rank order (identical bits on every
 * rank, run-to-run deterministic).  slots[p]: rank p's table of pmb_peer_reduce_table_doubles(world) doubles ([rank] = the
 * caller's own).  Replaces the NCCL all-reduce of the CG / LDAS dot products (pymoto/solvers/iterative.py:365-398 evaluates
 * them with numpy on one process). */
#define PMB_PEER_MAX 16
#define PMB_PEER_COUNT_MAX 16
typedef struct pmb_peer_reduce {
  int world, rank;
  double* slots[PMB_PEER_MAX];
  unsigned long long* ctl; /* 2 words of this rank's device memory: all-reduce counter, first timed-out all-reduce */
} pmb_peer_reduce;
long long pmb_peer_reduce_table_doubles(int world);
int pmb_peer_allreduce(const pmb_peer_reduce* r, double* val, int count, int op, void* stream);

/* ---- slab communication over NCCL: lets a host without torch.distributed drive the z-slab path.  NCCL is
 * bound at run time (dlopen libnccl.so.2); without it these return an error.  One handle per process / GPU, made from a
 * 128-byte ncclUniqueId that rank 0 creates (pmb_comm_unique_id) and the caller distributes.  z-neighbours are rank +- 1.
 * pmb_halo_exchange: base[own_offset, own_offset + own_len) are the owned doubles of a padded device buffer, the n doubles
 * below / above are halos; lower / upper select which of MY halos are refilled (all ranks pass the same flags).
 * pmb_allreduce: in-place sum of `count` device doubles (dot products, compliance, volume). */
typedef struct pmb_comm pmb_comm;
int pmb_comm_unique_id(void* id128);
int pmb_comm_init(const void* id128, int rank, int nranks, pmb_comm** out);
int pmb_comm_destroy(pmb_comm* c);
int pmb_comm_rank(const pmb_comm* c);
int pmb_comm_size(const pmb_comm* c);
int pmb_halo_exchange(pmb_comm* c, double* base, long long own_offset, long long own_len, long long n, int lower, int upper,
                      void* stream);
int pmb_allreduce(pmb_comm* c, double* buf, long long count, void* stream);

/* OC update, one bisection candidate (pymoto/common/optimizers.py:425-435): xnew = clip(x sqrt(-min(dg,0)/lmid),
 * max(xmin, x-move), min(xmax, x+move)), sum_out = sum(xnew) (deterministic). xnew may be NULL. ws: pmb_ws_doubles(). */
int pmb_oc_candidate(long long n, const double* x, const double* dg, double move, double xmin, double xmax, double lmid,
                     double* xnew, double* sum_out, double* ws, void* stream);

/* SIMP glue kept on device for the resident path: s = xmin + (1-xmin) y^p ; dy = ds * p (1-xmin) y^(p-1) */
int pmb_simp(long long n, double xmin, int p, const double* y, double* s, void* stream);
int pmb_simp_bwd(long long n, double xmin, int p, const double* y, const double* ds, double* dy, void* stream);

/* ---- MMA design update (pymoto/common/mma.py), n-sized parts; the m-sized unknowns stay with the caller on the host.
 * m = number of general constraints (1..PMB_MMA_MAXM; an unconstrained problem passes one dummy row of zeros like the
 * reference does).  P, Q: (m+1) x n row-major.  All `out` arrays are DEVICE memory, sums first then maxima; every pass is a
 * deterministic two-stage reduction over ws (pmb_mma_ws_doubles() doubles, zero-initialised once by the caller). */
#define PMB_MMA_MAXM 6
typedef struct {
  double s;        /* value for every variable ...                  */
  const double* v; /* ... unless v != NULL: per-variable device array */
} pmb_bound;
typedef struct {
  double *x, *xsi, *eta;           /* subproblem iterate: primal x and the multipliers of alfa <= x <= beta */
  double *xo, *xsio, *etao;        /* line-search base point (written by pmb_mma_newton_dir)                */
  double *dx, *dxsi, *deta;        /* Newton direction                                                       */
  double *low, *upp, *alfa, *beta; /* asymptotes, move-limited bounds                                        */
  double *P, *Q;                   /* (m+1) x n                                                              */
} pmb_mma_vecs;
long long pmb_mma_ws_doubles(void);
/* offset *= asyincr / asydecr by the sign of (x-xold1)(xold1-xold2), clipped to [1/asybound^2, asybound] (mma.py:129-140) */
int pmb_mma_asymptotes(long long n, const double* x, const double* xold1, const double* xold2, double asyincr, double asydecr,
                       double asybound, double* offset, void* stream);
/* mmasub set-up (mma.py:178-217) + subsolv start point (:288-293).  dg: HOST array of m+1 device row pointers; rho: HOST
 * array of m+1 values; version 1987 | 2007.  out[i] = sum_j (P_ij + Q_ij) / shift_j, i = 0..m (rhs = out - g on the host). */
int pmb_mma_setup(long long n, int m, const double* xval, const double* const* dg, const double* offset, pmb_bound xmin,
                  pmb_bound xmax, pmb_bound move, double albefa, const double* rho, int version, const pmb_mma_vecs* v,
                  double* out, double* ws, void* stream);
/* lam, dlam: HOST arrays of m values.  residual / linesearch: out = [sum of squared x-, xsi-, eta-residuals, gvec[m], max
 * squared residual]; newton_sums: out = [gvec[m], GG (delx/diagx) [m], (GG/diagx) GG^T [m*m]]; newton_dir stores dx, dxsi,
 * deta, xo, xsio, etao and out = [0, max(-dxsi/xsi), max(-deta/eta), max(-dx/(x-alfa)), max(dx/(beta-x))]; linesearch sets
 * (x, xsi, eta) = base + steg * direction before evaluating the residual. */
int pmb_mma_residual(long long n, int m, const pmb_mma_vecs* v, const double* lam, double epsi, double* out, double* ws, void* stream);
int pmb_mma_newton_sums(long long n, int m, const pmb_mma_vecs* v, const double* lam, double epsi, double* out, double* ws, void* stream);
int pmb_mma_newton_dir(long long n, int m, const pmb_mma_vecs* v, const double* lam, const double* dlam, double epsi, double* out,
                       double* ws, void* stream);
int pmb_mma_linesearch(long long n, int m, const pmb_mma_vecs* v, const double* lam, double steg, double epsi, double* out,
                       double* ws, void* stream);
/* GCMMA (mma.py:104-160, 232-242; the inner loop and the rho update stay with the caller on the host).  gcmma_rho:
 * out[i] = sum_j (xmax_j - xmin_j) |dg_i[j]|, i = 0..m (:151; rho = 0.1 / n * out; the subproblem is then set up with version 2007 and
 * max(rho_i, 1e-6), which is the reference's GCMMA P / Q formula, :213-216).  gcmma_estimate, after the subproblem solve:
 * out[i] = sum_j P_ij / (upp_j - x_j) + Q_ij / (x_j - low_j), i = 0..m (gest = out - rhs, :236) and out[m+1] = dk (:239). */
int pmb_mma_gcmma_rho(long long n, int m, const double* const* dg, pmb_bound xmin, pmb_bound xmax, double* out, double* ws, void* stream);
int pmb_mma_gcmma_estimate(long long n, int m, const pmb_mma_vecs* v, const double* xval, pmb_bound xmin, pmb_bound xmax, double* out,
                           double* ws, void* stream);

/* ---- whole linear solve driven from C (pymoto/solvers/iterative.py:340-403 CG.solve with :222-256 GeometricMultigrid.solve
 * as preconditioner): the same kernel launches as the entry points above, issued in the reference's order, the host
 * polling one scalar (the residual norm) per iteration.  Single GPU (every grid kz0 = 0, nzl = nz + 1).  All pointers
 * inside the descriptor are DEVICE memory owned by the caller except gen.Ke_host. */
#define PMB_MAX_LEVELS 12
typedef struct {
  pmb_grid grid;       /* this level                                                              */
  const double* A;     /* stencil-CSR values (may be NULL on level 0 when the generator is given) */
  const double* diag;  /* diagonal of A                                                           */
  double *u, *u2, *t;  /* scratch vectors, pmb_nrows(grid) doubles each                           */
  double* rc;          /* restricted residual, pmb_nrows(next coarser grid) doubles               */
  int smooth_steps;    /* damped-Jacobi sweeps before and after the coarse correction             */
  double w;            /* damping                                                                 */
} pmb_mg_level;
typedef struct {
  int nlevels;                        /* smoothed levels, finest first                                          */
  pmb_mg_level level[PMB_MAX_LEVELS];
  pmb_grid coarse_grid;               /* 2:1 coarsening of the last smoothed level, solved directly             */
  const double* coarse_inv;           /* its dense inverse (pmb_densify + pmb_dense_invert), row-major          */
  double* coarse_out;                 /* pmb_nrows(coarse_grid) doubles                                         */
  pmb_elem_op gen;                    /* level-0 generator for pmb_elem_spmv; gen.Ke_host == NULL: stream A      */
} pmb_mg_desc;
/* z = one V-cycle applied to r (z, r: pmb_nrows(level[0].grid) doubles) */
int pmb_vcycle(const pmb_mg_desc* mg, const double* r, double* z, void* stream);
/* Preconditioned CG for A x = b from the start vector in x (overwritten by the solution); r, q, p: n-double scratch; scal: 16
 * device doubles; ws_red: pmb_ws_doubles() zero-initialised doubles; ws_spmv: max(pmb_spmv_ws_doubles, pmb_elem_ws_doubles)
 * of level 0.  Stops when |r|/|b| <= tol or after maxit iterations; explicit residual every `restart` iterations.  *iters
 * = products A p, *relres = last |r|/|b| (both HOST). */
int pmb_pcg_solve(const pmb_mg_desc* mg, const double* b, double* x, double* r, double* q, double* p, double tol, int maxit,
                  int restart, double* scal, double* ws_red, double* ws_spmv, int* iters, double* relres, void* stream);

/* The same solve through a PLAN: descriptor + bound vectors (b: right-hand side, x: start vector in / solution out; r, q, p,
 * scal, ws_red, ws_spmv as for pmb_pcg_solve) + a CUDA graph of one CG iteration (V-cycle, direction update, product, x / r
 * update: everything between two host polls), captured on its second occurrence and replayed from then on, also by later
 * solves through the same plan.  Valid while the addresses in the descriptor, the element matrix behind gen.Ke_host and the
 * kernel layout stay the same (the operator VALUES may change: they are read from memory).  Iterates are bit-identical to
 * pmb_pcg_solve.  The plan is the only object the solver side of the library allocates; the caller destroys it. */
typedef struct pmb_pcg_plan pmb_pcg_plan;
int pmb_pcg_plan_create(const pmb_mg_desc* mg, const double* b, double* x, double* r, double* q, double* p, double* scal,
                        double* ws_red, double* ws_spmv, pmb_pcg_plan** plan);
int pmb_pcg_plan_solve(pmb_pcg_plan* plan, double tol, int maxit, int restart, int* iters, double* relres, void* stream);
long long pmb_pcg_plan_graph_replays(const pmb_pcg_plan* plan);
int pmb_pcg_plan_destroy(pmb_pcg_plan* plan);

/* In-run FP64 peak probe for bench.py's roofline (not on the product path): kind 0 = DFMA, 1 = DMMA.8x8x4 register-only
 * streams at 16 warps / SM; *tflops_out (HOST) = best of 3 launches; out: pmb_probe_fp64_out_doubles() device doubles. */
long long pmb_probe_fp64_out_doubles(void);
int pmb_probe_fp64(int kind, int iters, double* out, double* tflops_out, void* stream);

/* VTI writer payload: out[i*ncomp_out + c] = (float) in[i*ncomp_in + c], zero for c >= ncomp_in (round to nearest even) */
int pmb_pack_f32(long long nitems, int ncomp_in, int ncomp_out, const double* in, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
