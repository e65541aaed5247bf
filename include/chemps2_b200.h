/* chemps2_b200.h — C ABI of the B200-native two-site DMRG sweep hot path.
 *
 * The reference (SebWouters/CheMPS2) has no FFI seam for this path: the boundary is its C++ classes
 * (SURVEY.md 8(b)).  Every entry point below names the reference interface it replaces; INTEGRATION.md shows
 * the C++ shim a CheMPS2 maintainer would put behind the unchanged public headers.
 *
 * Conventions
 *   - plain C types only, opaque handles, caller-owned host buffers, library-owned device buffers;
 *   - every function returns 0 on success and a negative code on error; b2_last_error() gives the message
 *     (the reference asserts/aborts on impossible input: the C++ shim turns a non-zero code into abort());
 *   - packed tensor layouts are exactly the reference's gStorage() layouts (column-major blocks, enumeration
 *     orders of TensorT.cpp:38-104, TensorOperator.cpp:29-102, Sobject.cpp:36-147), so host arrays can be
 *     handed over unchanged;
 *   - all compute runs on the GPU; there is NO CPU fallback.  A context created with device = -1 is a
 *     planning-only context (sector tables, plans, exports); compute calls on it fail with B2_ERR_NO_DEVICE.
 */
#ifndef CHEMPS2_B200_H
#define CHEMPS2_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2_OK 0
#define B2_ERR_ARG (-1)
#define B2_ERR_NO_DEVICE (-2)
#define B2_ERR_CUDA (-3)
#define B2_ERR_STATE (-4)

/* operator kinds; (two_j, n_elec): L(1,1) S0(0,2) S1(2,2) F0(0,0) F1(2,0) A(0,2) B(2,2) C(0,0) D(2,0) Q(1,1) X(0,0) */
enum { B2_L = 0, B2_S0, B2_S1, B2_F0, B2_F1, B2_A, B2_B, B2_C, B2_D, B2_Q, B2_X,
       /* helper operators of the two-orbital correlation functions: G, Y, Z (0,0) TensorGYZ.cpp, K, M (1,1) TensorKM.cpp */
       B2_G, B2_Y, B2_Z, B2_K, B2_M };

typedef struct b2_ctx b2_ctx;
typedef struct b2_opset b2_opset;
typedef struct b2_heff b2_heff;
typedef struct b2_mps b2_mps;

const char* b2_last_error(void);
const char* b2_version(void);

/* ------------------------------------------------------------------------------------------------ context
 * device >= 0: CUDA ordinal (one process per GPU).  device = -1: planning-only (no GPU touched). */
int b2_ctx_create(int device, b2_ctx** out);
void b2_ctx_destroy(b2_ctx* ctx);
int b2_ctx_device(const b2_ctx* ctx);
/* run every kernel / copy of this context on the caller's CUDA stream (a cudaStream_t, e.g. torch's current stream) */
int b2_ctx_set_stream(b2_ctx* ctx, void* cuda_stream);
void* b2_ctx_stream(const b2_ctx* ctx);

/* ------------------------------------------------------------------------------------------------ problem
 * Replaces CheMPS2::Problem (Problem.h:44-128) as seen by the hot path: target sector, orbital irreps in DMRG
 * order and the dense table gMxElement(a,b,c,d) = mx[a + L*(b + L*(c + L*d))] (Problem.cpp:351-384).
 * b2_problem_set takes the folded table as the reference holds it; b2_problem_set_integrals builds it from
 * T (i + L*j) and physicist V <ab|cd> exactly like Problem::construct_mxelem. */
int b2_problem_set(b2_ctx* ctx, int L, int group, int N, int twoS, int irrep, const int* orb_irrep, const double* mx_elem,
                   double econst);
int b2_problem_set_integrals(b2_ctx* ctx, int L, int group, int N, int twoS, int irrep, const int* orb_irrep,
                             const double* tmat, const double* vmat, double econst);

/* Problem::setMxElement after the fact (Problem.cpp:357-361; the reference's tests/test12.cpp.in writes its model Hamiltonian this way):
 * replaces the folded table of the problem already set, the bookkeeper stays.  Operator sets and plans built before hold the old
 * integrals: rebuild them (b2_dmrg_presolve drops the sweep driver's cached plans). */
int b2_problem_update_mx(b2_ctx* ctx, const double* mx_elem);
/* copy of the folded table gMxElement (L^4 doubles) as the library holds it */
int b2_problem_mx(const b2_ctx* ctx, double* mx_out);
/* Wigner 6j / 9j symbols with doubled arguments, as Wigner::wigner6j / wigner9j (Wigner.cpp:294-368) */
double b2_wigner6j(int two_ja, int two_jb, int two_jc, int two_jd, int two_je, int two_jf);
double b2_wigner9j(int two_ja, int two_jb, int two_jc, int two_jd, int two_je, int two_jf, int two_jg, int two_jh, int two_ji);

/* eigen-decomposition of a small symmetric matrix (n <= 32, column-major, ld = n), eigenvalues ascending: the host-side
 * Rayleigh-Ritz step of the device Davidson (stands in for dsyev_ at Davidson.cpp:276) */
int b2_small_symmetric_eig(int n, const double* a, double* eval, double* evec);

/* ------------------------------------------------------------------------------------------------ bookkeeper
 * Replaces CheMPS2::SyBookkeeper (SyBookkeeper.h:41-143).  b2_bk_init = constructor (FCI dims, ceil-scaled to D);
 * b2_bk_set_dim = SetDim; the getters mirror gCurrentDim / gFCIdim / gNmin / gNmax / gTwoSmin / gTwoSmax. */
int b2_bk_init(b2_ctx* ctx, int D);
int b2_bk_set_dim(b2_ctx* ctx, int boundary, int N, int twoS, int irrep, int dim);
int b2_bk_dim(const b2_ctx* ctx, int boundary, int N, int twoS, int irrep);
int b2_bk_fcidim(const b2_ctx* ctx, int boundary, int N, int twoS, int irrep);
int b2_bk_nmin(const b2_ctx* ctx, int boundary);
int b2_bk_nmax(const b2_ctx* ctx, int boundary);
int b2_bk_twosmin(const b2_ctx* ctx, int boundary, int N);
int b2_bk_twosmax(const b2_ctx* ctx, int boundary, int N);

/* packed sizes / block tables (for host code that needs the reference's kappa2index) */
int64_t b2_tensor_t_size(const b2_ctx* ctx, int site);                       /* TensorT::gKappa2index(gNKappa()) */
int64_t b2_sobject_size(const b2_ctx* ctx, int site);                        /* Sobject::gKappa2index(gNKappa()) */
int b2_sobject_nkappa(const b2_ctx* ctx, int site);
/* labels[9*k..]: NL,2SL,IL,N1,N2,2J,NR,2SR,IR ; offsets[k], k < nkappa+1 */
int b2_sobject_table(const b2_ctx* ctx, int site, int* labels, int64_t* offsets);

/* ------------------------------------------------------------------------------------------------ operator sets
 * One arena (device + host mirror) holding every renormalized operator of one boundary and one direction; replaces
 * the pointer tables Ltensors/F0tensors/.../Xtensors[boundary-1] of DMRG.h:211-226 allocated by
 * DMRG::allocateTensors (DMRGoperators.cpp:909-1145).  moving_right != 0: operators of the block left of `boundary`. */
int b2_opset_create(b2_ctx* ctx, int boundary, int moving_right, b2_opset** out);
/* the set {G, Y, Z, K, M}(site s) for every s < boundary, as DMRG::update_correlations_tensors keeps it
 * (DMRGoperators3RDM.cpp:415-479; moving right only, 1 <= boundary <= L-1).  b2_update_create on two such sets (old at boundary
 * index, new at index+1; old = NULL for index 0) replaces update_correlations_tensors(index+1): TensorGYZ::construct /
 * TensorKM::construct for the newest site, TensorOperator::update without Jordan-Wigner phase for the others. */
int b2_opset_create_correlation(b2_ctx* ctx, int boundary, b2_opset** out);
void b2_opset_destroy(b2_opset* set);
int b2_opset_count(const b2_opset* set);
int b2_opset_info(const b2_opset* set, int index, int* kind, int* site_i, int* site_j, int64_t* size);
int b2_opset_find(const b2_opset* set, int kind, int site_i, int site_j);   /* index or -1 */
/* packed data in the reference's TensorOperator storage order <-> gStorage() */
int b2_opset_upload(b2_opset* set, int index, const double* packed);
int b2_opset_download(b2_opset* set, int index, double* packed);
int b2_opset_clear(b2_opset* set);
/* life-cycle: move the arena of a set to pinned host memory and release its HBM / bring it back.  Replaces DMRG::OperatorsOnDisk,
 * deleteTensors, allocateTensors (DMRGoperators.cpp:33-231, 1147-1433; the reference spills to HDF5 files).  Compute entry points
 * refuse offloaded sets (B2_ERR_STATE); upload / download keep working on the host copy. */
int b2_opset_offload(b2_opset* set);
/* second tier: park the arena in a FILE (NVMe scratch, like the reference's CheMPS2_Operators_*.h5 files); neither HBM nor host memory is
 * held meanwhile, b2_opset_reload reads it back and removes the file */
int b2_opset_offload_file(b2_opset* set, const char* path);
int b2_opset_reload(b2_opset* set);
int b2_opset_resident(const b2_opset* set);
/* synthetic contents: element e of operator (kind,i,j) = amp * hash(seed, side, kind, i, j, e) in [-amp/2, amp/2); the
 * identical fill is produced by oracle/ref_driver.cpp `synth`, which lets bench.py compare with the reference at full size */
int b2_opset_fill_hash(b2_opset* set, uint64_t seed, double amp);
int b2_hash_fill(double* out, int64_t n, uint64_t seed, uint64_t key, double amp);

/* ------------------------------------------------------------------------------------------------ effective Hamiltonian
 * b2_heff_create builds the SigmaPlan for the site pair (site, site+1) from the operator sets at boundaries
 * `site` (left, may be NULL at the left edge) and `site+2` (right, may be NULL at the right edge).
 * world/rank: GPU sharding by the reference's ownership maps (MPIchemps2.h:158-231); world = 1 keeps every term.
 *
 * b2_heff_apply        = Heff::makeHeff (Heff.cpp:43-248): vec_out = H_eff * vec_in, both HOST buffers of
 *                        b2_heff_veclength doubles in the symmetric convention (Sobject.cpp:624-636).
 * b2_heff_apply_device = same on device-resident vectors (what the device Davidson uses).
 * b2_heff_diag         = Heff::fillHeffDiag (Heff.cpp:250-315). */
int b2_heff_create(b2_ctx* ctx, int site, b2_opset* left, b2_opset* right, int world, int rank, b2_heff** out);
void b2_heff_destroy(b2_heff* h);
int64_t b2_heff_veclength(const b2_heff* h);
int b2_heff_apply(b2_heff* h, const double* vec_in, double* vec_out);
int b2_heff_apply_device(b2_heff* h, const double* dev_in, double* dev_out);
int b2_heff_diag(b2_heff* h, double* diag);
int b2_heff_diag_device(b2_heff* h, double* dev_diag);
/* Excited states: the nLower / VeffTilde arguments of Heff::makeHeff / fillHeffDiag / SolveDAVIDSON (Heff.h:70).  veff_tilde[s] =
 * level-shifted lower state s projected on this site pair (HOST, veclength doubles, symmetric convention).  Afterwards every
 * apply adds sum_s <V_s|S> V_s (addDiagramExcitations, HeffDiagrams1.cpp:65-85) and the diagonal adds V_s .* V_s
 * (addDiagonalExcitations, HeffDiagonal.cpp:621-640).  n_lower = 0 switches it off. */
int b2_heff_set_excitations(b2_heff* h, int n_lower, const double* const* veff_tilde);
/* b2_heff_solve = Heff::SolveDAVIDSON (Heff.cpp:317-386) with CheMPS2::Davidson (Davidson.cpp) running on the device:
 * s (HOST, Sobject storage in the reference's "program" convention) holds the initial guess on entry and the lowest
 * eigenvector on exit; *eigenvalue excludes Econst exactly like the reference's return value; rtol = the sweep
 * instruction's Davidson tolerance (ConvergenceScheme); constants 32 / 3 / 1e-12 are Options.h:70-72.
 * b2_heff_solve_device = same with s resident on the device. */
int b2_heff_solve(b2_heff* h, double* s, double rtol, double* eigenvalue, int* n_matvec);
int b2_heff_solve_device(b2_heff* h, double* dev_s, double rtol, double* eigenvalue, int* n_matvec);
/* ------------------------------------------------------------------------------------------------ Davidson (reverse communication)
 * CheMPS2::Davidson (Davidson.h:46-58) as a device-backed object: b2_davidson_create = the constructor (veclength, MAX_NUM_VEC, NUM_VEC_KEEP,
 * RTOL, DIAG_CUTOFF; problem type 'E'), b2_davidson_fetch = FetchInstruction, b2_davidson_num_multiplications = GetNumMultiplications.
 * The pointers handed out are DEVICE pointers (veclength doubles each) valid until the next fetch:
 *   'A'  write the initial guess into ptr0 and the diagonal of the matrix into ptr1, then fetch again
 *   'B'  compute ptr1 = H * ptr0 on the context stream (e.g. b2_heff_apply_device), then fetch again
 *   'C'  converged: ptr0 = the lowest eigenvector (unit norm), ptr1[0] = its eigenvalue (also b2_davidson_eigenvalue)
 * The algorithm is the one b2_heff_solve runs (Options.h:70-72 constants there); b2_heff_solve is this loop fused with the sigma build. */
typedef struct b2_davidson b2_davidson;
int b2_davidson_create(b2_ctx* ctx, int64_t veclength, int max_num_vec, int num_vec_keep, double rtol, double diag_cutoff, b2_davidson** out);
void b2_davidson_destroy(b2_davidson* d);
int b2_davidson_fetch(b2_davidson* d, char* instruction, double** dev_ptr0, double** dev_ptr1);
int b2_davidson_num_multiplications(const b2_davidson* d);
double b2_davidson_eigenvalue(const b2_davidson* d);
/* multi-GPU: callback that sums a device vector over all ranks in place on the given stream (the caller owns the NCCL
 * communicator); replaces MPI_Reduce/MPI_Bcast of Heff.cpp:350-365.  Return 0 on success. */
typedef int (*b2_allreduce_fn)(void* user, double* dev_ptr, int64_t n, void* cuda_stream);
int b2_heff_set_allreduce(b2_heff* h, b2_allreduce_fn fn, void* user);
/* statistics: [0] #terms, [1] #terms dropped (zero prefactor), [2] #presummed operators, [3] reference FLOPs per apply
 * (2mnk per reference dgemm_), [4] executed FLOPs per apply, [5] workspace doubles, [6] #stage-1 GEMMs, [7] #CTAs,
 * [8] #waves, [9] kernel launches per apply, [10] split-K partial doubles, [11] bytes of device work lists */
int b2_heff_stats(const b2_heff* h, double* out12);
/* seconds spent in the kernels of the last b2_heff_apply* call, measured with CUDA events on the launch stream */
double b2_heff_last_kernel_seconds(const b2_heff* h);

/* flat export of the plan (tests / the CPU checker in oracle/ only) */
typedef struct {
   int32_t dst, src;          /* Sobject block ids */
   int32_t a_rows, a_cols;    /* stored shape of the left operator block (0 when absent) */
   int32_t b_rows, b_cols;    /* stored shape of the right operator block */
   int8_t a_space, a_trans;   /* 0 none, 1 left arena, 2 right arena, 3 presum arena */
   int8_t b_space, b_trans;
   int32_t owner;
   int64_t a_off, b_off;      /* offsets (doubles) inside the arena named by *_space */
   double factor;
} b2_flat_term;
typedef struct {
   int64_t dst_off;           /* in the presum arena */
   int64_t src_off;           /* in the arena of `space` */
   int64_t size;
   int32_t space;             /* 1 left arena, 2 right arena */
   double coef;
} b2_flat_presum;
int64_t b2_heff_num_terms(const b2_heff* h);
int b2_heff_export_terms(const b2_heff* h, b2_flat_term* out);
int64_t b2_heff_num_presum_parts(const b2_heff* h);
int64_t b2_heff_presum_size(const b2_heff* h);
int b2_heff_export_presums(const b2_heff* h, b2_flat_presum* out);
/* FP64 peak probe, register resident: use_mma = 1 times DMMA (mma.sync m8n8k4 f64), 0 times DFMA; result in TFLOP/s */
int b2_probe_fp64(b2_ctx* ctx, int use_mma, double* tflops);
/* raw view of the compiled device work lists (structs of chemps2_b200/csrc/b2_device.h / b2_heff.h); used by the
 * work-list emulator in oracle/ that checks the scheduling (waves, split-K, reduces) on the CPU */
typedef struct {
   const void *items1, *items2, *tiles1[4], *tiles2[4], *reduces, *waves;
   int64_t n_items1, n_items2, n_tiles1[4], n_tiles2[4], n_reduces, n_waves, work_size, part_size;
} b2_worklists;
int b2_heff_worklists(const b2_heff* h, b2_worklists* out);
/* the lists behind b2_heff_diag (DiagItem / DiagTile of chemps2_b200/csrc/b2_device.h: diag[tile](i, j) = sum_items f * A(i,i) * B(j,j),
 * the terms of Heff::fillHeffDiag, Heff.cpp:250-315 + HeffDiagonal.cpp) for the CPU checker */
int b2_heff_diag_lists(const b2_heff* h, const void** items, int64_t* n_items, const void** tiles, int64_t* n_tiles);

/* ------------------------------------------------------------------------------------------------ operator update
 * Replaces DMRG::updateMovingRight / updateMovingLeft (DMRGoperators.cpp:243-907) together with the tensor algebra they
 * call: TensorOperator::update (TensorOperator.cpp:163-405), TensorL::create (TensorL.cpp:41-206), TensorS0/S1/F0/F1::makenew,
 * the A/B/C/D daxpy mixing (DMRGoperators.cpp:367-405), TensorQ::AddTerm* and TensorX::update.
 * index = site of the MPS tensor T that was just optimised.  moving_right != 0: old_set sits at boundary `index` (NULL when
 * index == 0), new_set at boundary index+1.  moving_right == 0: old_set at boundary index+1 (NULL when index == L-1), new_set
 * at boundary `index`.  b2_update_run takes T as the reference's TensorT::gStorage() (HOST); every operator of new_set is
 * overwritten on the device (b2_opset_download brings one back). */
typedef struct b2_update b2_update;
int b2_update_create(b2_ctx* ctx, int index, int moving_right, b2_opset* old_set, b2_opset* new_set, b2_update** out);
void b2_update_destroy(b2_update* u);
/* multi-GPU: the NEW operators are distributed over `world` GPUs (deterministic FLOP-balanced assignment; replaces the static
 * owner maps of MPIchemps2.h:158-231 for the update, whose Q/X partial exchanges are DMRGoperators.cpp:449-533); rank `rank`
 * computes its share, the all-reduce callback sums the arenas so that every GPU ends up with every operator. */
int b2_update_create_sharded(b2_ctx* ctx, int index, int moving_right, b2_opset* old_set, b2_opset* new_set, int world, int rank,
                             b2_update** out);
int b2_update_set_allreduce(b2_update* u, b2_allreduce_fn fn, void* user);
int b2_update_run(b2_update* u, const double* t_host);
int b2_update_run_device(b2_update* u, const double* t_dev);
/* [0] #terms, [1] #mix terms, [2] #presums, [3] reference FLOPs, [4] executed FLOPs, [5] workspace doubles, [6] #waves, [7] launches */
int b2_update_stats(const b2_update* u, double* out8);
/* work lists of pass 0 (contractions) / pass 1 (transposed copies for the A,B,C,D mixing, written into the pre-sum arena) and the pre-sum
 * jobs, for the CPU emulator in oracle/ */
int b2_update_worklists(const b2_update* u, int pass, b2_worklists* out);
/* the whole-operator axpys of the A/B/C/D mixing (dst in the new arena += coef * src; space 6 = new arena, 3 = pre-sum arena, where pass 1
 * has put the transposed copies daxpy_transpose_tensorCD needs), run after pass 1 — for the CPU emulator */
int64_t b2_update_num_mix_flat(const b2_update* u);
int b2_update_export_mix_flat(const b2_update* u, b2_flat_presum* out);
int64_t b2_update_num_presum_parts(const b2_update* u);
int64_t b2_update_presum_size(const b2_update* u);
int b2_update_export_presums(const b2_update* u, b2_flat_presum* out);

/* ------------------------------------------------------------------------------------------------ sweep driver
 * The part of CheMPS2::DMRG the hot path lives in (DMRG.cpp:357-452): it owns the MPS site tensors (host, TensorT::gStorage()
 * layouts) and one operator set per boundary and direction (device), and strings the pieces together:
 *   b2_dmrg_update      = DMRG::updateMovingRight(index) (moving_right != 0: operators of boundary index+1 from MPS[index])
 *                         / updateMovingLeft (moving_right == 0: operators of boundary index from MPS[index])
 *   b2_dmrg_solve_site  = DMRG::solve_site: Sobject::Join (device) -> Heff::SolveDAVIDSON (device) -> addNoise -> Sobject::Split
 *                         (host SVD; virtual dimensions of boundary index+1 are rewritten when change != 0); *energy includes Econst
 *   b2_dmrg_sweep       = DMRG::sweepleft (to_right == 0: index L-2 .. 1) / sweepright (index 0 .. L-3) incl. the operator updates;
 *                         `noise` is the ConvergenceScheme noise PREFACTOR: like DMRG.cpp:360,391 the level handed to solve_site is
 *                         |noise| x the largest discarded weight of the previous half sweep (0 before the first one);
 *                         b2_dmrg_solve_site takes the absolute level, exactly like DMRG::solve_site */
typedef struct b2_dmrg b2_dmrg;
int b2_dmrg_create(b2_ctx* ctx, b2_dmrg** out);
void b2_dmrg_destroy(b2_dmrg* d);
int64_t b2_dmrg_mps_size(const b2_dmrg* d, int site);
int b2_dmrg_set_mps(b2_dmrg* d, int site, const double* t_storage);
int b2_dmrg_get_mps(const b2_dmrg* d, int site, double* t_storage);
/* Random MPS exactly as DMRG::setupBookkeeperAndMPS builds it after the caller's srand(seed) (DMRG.cpp:149-169): TensorT::random
 * (TensorT.cpp:167-173) site by site from the stream of glibc's rand(), then left-normalisation with LAPACK's Householder conventions
 * (TensorT::QR, TensorT.cpp:188-265) — the tensors equal the reference's.  The same stream later feeds Sobject::addNoise
 * (Sobject.cpp:652-659) in every b2_dmrg_solve_site with noise > 0, so seeded noisy sweeps follow the reference step by step.
 * b2_dmrg_srand only re-seeds that stream (a run that starts from a checkpoint); b2_rand_stream exposes it (out[i] = i-th rand()). */
int b2_dmrg_random_mps(b2_dmrg* d, uint64_t seed);
int b2_dmrg_srand(b2_dmrg* d, uint64_t seed);
int b2_rand_stream(uint64_t seed, int n, int* out);
b2_opset* b2_dmrg_opset(b2_dmrg* d, int boundary, int moving_right);
int b2_dmrg_set_opset(b2_dmrg* d, int boundary, int moving_right, b2_opset* set);   /* the driver takes ownership */
/* MPS checkpoint (DMRG::saveMPS / loadDIM / loadMPS, DMRGmpsio.cpp:30-131): converged flag, all virtual dimensions, the TensorT storage of
 * every site in one flat binary file (no HDF5 in this image: same payload, different container).  Loading re-dimensions the bookkeeper,
 * replaces the MPS and drops every operator set (run b2_dmrg_presolve again). */
int b2_dmrg_save_mps(const b2_dmrg* d, const char* path, int converged);
int b2_dmrg_load_mps(b2_dmrg* d, const char* path, int* converged);
/* b2_dmrg_presolve = DMRG::PreSolve (DMRG.cpp:257-266).  b2_dmrg_solve = DMRG::Solve (DMRG.cpp:268-355) with the ConvergenceScheme
 * (ConvergenceScheme.h) passed as arrays of length n_instructions; *energy = lowest energy encountered (Econst included). */
int b2_dmrg_presolve(b2_dmrg* d);
int b2_dmrg_solve(b2_dmrg* d, int n_instructions, const int* D, const double* energy_conv, const int* max_sweeps, const double* noise_prefactor,
                  const double* davidson_rtol, double* energy);
/* Excited states (DMRG::activateExcitations / newExcitation, DMRG.cpp:464-505): the current MPS is stored as a lower state with level shift
 * `eshift`; a fresh random MPS (bookkeeper re-initialised like b2_bk_init(D)) takes its place and all operator sets are dropped — run
 * the PreSolve updates again, then sweep: every b2_dmrg_update also renews the overlap tensors with the stored states (TensorO,
 * DMRGoperators.cpp:556-567, 889-900) and every b2_dmrg_solve_site adds the projector sum_s Eshift_s |s><s| through calcVeffTilde
 * (DMRGtechnics.cpp:540-620) + Heff::addDiagramExcitations. */
int b2_dmrg_new_excitation(b2_dmrg* d, double eshift, int D, uint64_t seed);
int b2_dmrg_num_lower_states(const b2_dmrg* d);
/* 2-RDM of the current MPS: the TwoDM part of DMRG::calc_rdms_and_correlations (DMRGtechnics.cpp:40-113) — MPS into left-canonical form
 * (device-SVD gauge moves), operators of every boundary, then from the right site by site b2_twodm_fill_site, right-normalise, next
 * moving-left operators; finally TwoDM::correct_higher_multiplicities.  two_rdm_A / two_rdm_B: L^4 doubles each, DMRG orbital order,
 * A(i,j,k,l) at i + L*(j + L*(k + L*l)) as TwoDM::getTwoDMA_DMRG.  The MPS is left in right-canonical form with its norm set to 1.
 * The chain keeps REDUCED operator sets (L only on the left, L/S0/S1/F0/F1 on the right, like updateMovingLeftSafe2DM); a sigma
 * build refuses them (B2_ERR_STATE): rebuild the sweep operators with b2_dmrg_update before sweeping again. */
int b2_dmrg_calc_2rdm(b2_dmrg* d, double* two_rdm_A, double* two_rdm_B);
/* Correlations (Correlations.cpp; the second half of DMRG::calc_rdms_and_correlations, DMRGtechnics.cpp:150-175): from the finished
 * 2-RDM arrays A, B (b2_dmrg_calc_2rdm) the L x L tables Cspin, Cdens, Cspinflip, Cdirad (FillSpinDensSpinflip, :69-103) and the
 * two-orbital mutual information MutInfo (Correlations::FillSite, :212-351) of the current MPS; element (row, col) at row + L*col as
 * Correlations::getCspin_DMRG etc.  b2_corr_fill_site is the per-site step (T = MPS[site] as orthogonality centre, corr = the
 * correlation operator set of boundary `site`). */
int b2_dmrg_calc_correlations(b2_dmrg* d, const double* two_rdm_A, const double* two_rdm_B, double* Cspin, double* Cdens, double* Cspinflip,
                              double* Cdirad, double* MutInfo);
int b2_corr_fill_site(b2_ctx* ctx, int site, const double* t_host, b2_opset* corr, const double* two_rdm_A, const double* two_rdm_B, double* Cdirad,
                      double* MutInfo);
/* multi-GPU sweep: sigma terms (ownership maps) and operator updates are sharded over `world` GPUs, MPS / Davidson vectors /
 * Split are replicated; fn sums a device vector over the ranks (NCCL).  Call before the first update / solve. */
int b2_dmrg_set_world(b2_dmrg* d, int world, int rank, b2_allreduce_fn fn, void* user);
/* enabled != 0: only the two operator sets of the site pair being optimised (and the set being built) stay in HBM, every
 * other boundary is offloaded to pinned host memory (the reference's disk mode, DMRG.cpp:57-65 makecheckpoints / OperatorsOnDisk) */
int b2_dmrg_set_spill(b2_dmrg* d, int enabled);
/* dir != NULL / non-empty: spill mode parks the operator sets in files of that directory (b2_opset_offload_file) instead of pinned host
 * memory — the reference's tmp folder of DMRG::DMRG(..., tmpfolder) */
int b2_dmrg_set_spill_dir(b2_dmrg* d, const char* dir);
/* The driver keeps the sigma plan of the last visit of every site (device work lists only) and re-uses it when the dimension tables of
 * the two boundaries the plan reads (site and site + 2; the contracted one in between is re-dimensioned by every visit and not part of
 * the two-site object, Sobject.cpp:36-78) are unchanged — the normal situation in converged sweeps at a fixed virtual dimension.
 * enabled = 0 switches the cache off and frees it; the statistics count re-used and newly built plans. */
int b2_dmrg_set_plan_cache(b2_dmrg* d, int enabled);
int b2_dmrg_plan_cache_stats(const b2_dmrg* d, long long* hits, long long* misses);
/* Inside b2_dmrg_sweep the host half of the NEXT site's sigma plan (term enumeration + scheduling; it needs the dimension tables and the
 * layouts of the operator sets, not their contents) is built on a helper thread while the calling thread plans and runs the operator
 * update that precedes it (DMRG.cpp:372-377 runs the two one after the other).  enabled = 0 switches it off (also B2_PLAN_PREFETCH=0);
 * the result of a sweep does not depend on it.  Sharded sweeps (b2_dmrg_set_world with world > 1) use it only with enabled = 2 (or
 * B2_PLAN_PREFETCH=2): it has been validated on one GPU only.  b2_dmrg_plan_prefetched: how many newly built plans came from the helper thread. */
int b2_dmrg_set_plan_prefetch(b2_dmrg* d, int enabled);
long long b2_dmrg_plan_prefetched(const b2_dmrg* d);
/* wall-clock seconds per phase since the last reset: [0] plan building (host), [1] Davidson solves, [2] Split (host SVD),
 * [3] operator updates, [4] number of sigma builds */
int b2_dmrg_timers(b2_dmrg* d, double* out5, int reset);
int b2_dmrg_update(b2_dmrg* d, int index, int moving_right);
int b2_dmrg_solve_site(b2_dmrg* d, int index, double rtol, double noise, int D, int moving_right, int change, double* energy,
                       double* discarded_weight, int* n_matvec);
int b2_dmrg_sweep(b2_dmrg* d, int to_right, double rtol, double noise, int D, int change, double* min_energy, double* max_discarded);
/* the sweep counters DMRG keeps (DMRG.cpp:356-414): [0] energy of the last site solved (the value sweepleft / sweepright return and
 * DMRG::Solve tests for convergence), [1] LastMinEnergy, [2] MaxDiscWeightLastSweep, [3] TotalMinEnergy since the last PreSolve */
int b2_dmrg_sweep_info(const b2_dmrg* d, double* out4);

/* ------------------------------------------------------------------------------------------------ Sobject::Join
 * b2_join_* = Sobject::Join (Sobject.cpp:212-258): the two-site object of sites (site, site+1) from their site tensors,
 *    S[kappa] = sum_jM phase * sqrt((2J+1)(2jM+1)) * 6j * T_site[L -> M] * T_site+1[M -> R]      (program convention, as Sobject::gStorage()).
 * The plan holds the three-factor terms for the current bookkeeper dimensions; b2_join_run takes TensorT::gStorage() of both sites as
 * host buffers and fills s_out (b2_sobject_size doubles) on the GPU; b2_join_worklists exports the compiled lists for the CPU checker. */
typedef struct b2_join b2_join;
int b2_join_create(b2_ctx* ctx, int site, b2_join** out);
void b2_join_destroy(b2_join* j);
int b2_join_run(b2_join* j, const double* t_left, const double* t_right, double* s_out);
int b2_join_worklists(const b2_join* j, b2_worklists* out);

/* b2_sobject_split = Sobject::Split (Sobject.cpp:260-622): s_storage (program convention, the layout of b2_sobject_table for the CURRENT
 * dimensions) is recoupled into one matrix per centre sector, decomposed, truncated to at most D states over all sectors by the
 * reference's global rule (:451-486, only when change != 0; the virtual dimensions of boundary site+1 in the bookkeeper are rewritten)
 * and scattered into the two new site tensors: moving_right != 0 -> left tensor left-normalised, the weights go right; else the
 * mirror image.  svd = NULL: all decompositions run together on the GPU (b2_svd_batch); otherwise the caller's routine is used (same
 * argument convention as b2_svd_batch, return 0 on success) - e.g. LAPACK dgesdd_ in a host program that keeps Split on the CPU.
 * The new tensors (sizes follow the NEW dimensions: b2_split_size) are fetched from the result handle. */
typedef struct b2_split b2_split;
typedef int (*b2_svd_fn)(void* user, int count, const int* m, const int* n, const double* const* a, double* const* s, double* const* u,
                         double* const* vt);
int b2_sobject_split(b2_ctx* ctx, int site, const double* s_storage, int D, int moving_right, int change, b2_svd_fn svd, void* user,
                     b2_split** out, double* discarded_weight);
int64_t b2_split_size(const b2_split* r, int right);          /* doubles of the new left (right = 0) / right (right != 0) site tensor */
int b2_split_get(const b2_split* r, int right, double* t_out);
void b2_split_destroy(b2_split* r);

/* ------------------------------------------------------------------------------------------------ 2-RDM
 * b2_twodm_fill_site = TwoDM::FillSite (TwoDM.cpp:445-628 with its 24 diagram functions doD1..doD24, :642-1592): the entries of the
 * spin-summed 2-RDM arrays two_rdm_A / two_rdm_B (L^4 doubles each, index c1 + L*(c2 + L*(c3 + L*c4)), DMRG orbital order, the four
 * symmetry-equivalent positions of set_2rdm_A_DMRG written together) that involve orbital `site` as the orthogonality centre.
 * t_host = TensorT::gStorage() of MPS[site]; left = operator set of boundary `site` moving right (NULL for site 0), right = operator
 * set of boundary site+1 moving left (NULL for site L-1); only their L / S0 / S1 / F0 / F1 operators are read.  The chain around it
 * (left/right normalisation, updateMovingLeftSafe2DM, correct_higher_multiplicities: DMRGtechnics.cpp:40-113) is host-side glue. */
int b2_twodm_fill_site(b2_ctx* ctx, int site, const double* t_host, b2_opset* left, b2_opset* right, double* two_rdm_A, double* two_rdm_B);
/* the same in two steps (plan once, run per tensor); planning works on planning-only contexts */
typedef struct b2_twodm b2_twodm;
int b2_twodm_create(b2_ctx* ctx, int site, b2_opset* left, b2_opset* right, b2_twodm** out);
void b2_twodm_destroy(b2_twodm* p);
int b2_twodm_run(b2_twodm* p, const double* t_host, double* two_rdm_A, double* two_rdm_B);
/* inspection for the CPU checker in oracle/ (tests only): work lists that build the effective operators (spaces: LEFT = left
 * operator arena, RIGHT = T, VOUT = arena of b2_twodm_m_size doubles); the groups of effective operators (dense [stride x members]
 * matrices at `off`, op_size valid doubles per column) with the operators of the left/right set they are paired with; the per-block
 * weights of diagram 1; and the scatter of externally computed inner products gram[group][member + members*partner] into A and B */
int b2_twodm_worklists(const b2_twodm* p, b2_worklists* out);
int64_t b2_twodm_m_size(const b2_twodm* p);
int b2_twodm_num_groups(const b2_twodm* p);
int b2_twodm_group_info(const b2_twodm* p, int g, int* left_side, int64_t* off, int64_t* stride, int64_t* op_size, int* n_members, int* n_partners,
                        int* partners, int cap);
int b2_twodm_d1_scale(const b2_twodm* p, double* per_block, int cap);
int b2_twodm_scatter(const b2_twodm* p, const double* const* gram, double d1, double* two_rdm_A, double* two_rdm_B);

/* Batched thin SVD on the GPU: a[i] (HOST, column-major m[i] x n[i], ld = m[i]) = U diag(s) V^T with k = min(m, n); u[i] is m x k
 * (ld m), vt[i] is k x n (ld k), s[i] decreasing.  Stands for the dgesdd_ call per centre sector of Sobject::Split
 * (Sobject.cpp:412-419); one-sided Jacobi, all matrices of the batch progress together (b2_svd.cu).  b2_dmrg_solve_site uses it. */
int b2_svd_batch(b2_ctx* ctx, int count, const int* m, const int* n, const double* const* a, double* const* s, double* const* u,
                 double* const* vt);

/* scheduling knobs of plans created afterwards: "work_budget" (doubles of stage-1 workspace per wave), "chunk_k",
 * "parallel_plan_flops" (plans below this many reference FLOPs per apply are compiled on all host cores; env B2_PLAN_THREADS),
 * "davidson_max_matvec" (safety net of the device Davidson: a solve that needs more matrix-vector products fails, default 5000) */
int b2_ctx_set_option(b2_ctx* ctx, const char* name, double value);
/* host mirrors of the operator arenas (valid until the set is destroyed) */
const double* b2_opset_host_arena(const b2_opset* set);
int64_t b2_opset_arena_size(const b2_opset* set);

#ifdef __cplusplus
}
#endif
#endif
