// chemps2_b200_shim.cpp — the drop-in: strong definitions of the hot-path member functions of CheMPS2 over the C ABI.
//
// This translation unit is compiled against the UNMODIFIED reference headers (byte-identical public AND private sections) and linked
// together with the UNMODIFIED reference sources into libchemps2.so.3 (dropin/build_dropin.sh).  The reference's own definitions of
//     Heff::SolveDAVIDSON   (Heff.cpp:317-329)      Heff::makeHeff (Heff.cpp:43-248)      Heff::fillHeffDiag (Heff.cpp:250-315)
//     DMRG::updateMovingRight (DMRGoperators.cpp:243-574)      DMRG::updateMovingLeft (DMRGoperators.cpp:576-907)
// are demoted to weak symbols in their object files (objcopy --weaken-symbol), so the definitions below win at link time and every
// caller inside the library — DMRG::solve_site, the updateMoving*Safe* wrappers, PreSolve, the 2-RDM / correlation / Fock chains, CASSCF —
// as well as the chemps2 binary, PyCheMPS2 and the reference's tests run their sigma builds, Davidson solves and operator updates on
// the GPU.  Everything else (Hamiltonian, Problem, SyBookkeeper, Sobject::Join/Split, TwoDM, HDF5 I/O, ...) is the reference's code.
//
// Host-mirror rule: the reference's tensors stay the owners of the host copies.  After every operator update the new
// operators are brought back into the TensorOperator storage the caller allocated (gStorage()), so unmodified consumers (TwoDM::FillSite,
// OperatorsOnDisk, Correlations, ...) keep working; the device copy is kept in a registry keyed by the DMRG object's table slot, and a
// sigma build whose tensors are the ones registered (same TensorX object, same content fingerprint) skips the upload.
//
// There is no CPU fallback: when the CUDA library reports an error the shim prints it and aborts, like the reference aborts on
// impossible input (its asserts).
#include <sys/time.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <map>
#include <vector>

#include "DMRG.h"
#include "Heff.h"
#include "Problem.h"
#include "Sobject.h"
#include "SyBookkeeper.h"
#include "TensorF0.h"
#include "TensorF1.h"
#include "TensorL.h"
#include "TensorO.h"
#include "TensorOperator.h"
#include "TensorQ.h"
#include "TensorS0.h"
#include "TensorS1.h"
#include "TensorT.h"
#include "TensorX.h"
#include "chemps2_b200.h"

namespace {

using namespace CheMPS2;

void die(const char* where, int rc) {
   std::fprintf(stderr, "chemps2_b200 drop-in: %s failed (%d): %s\n", where, rc, b2_last_error());
   std::abort();
}
#define B2(call) do { const int rc_ = (call); if (rc_ != B2_OK) die(#call, rc_); } while (0)

double wall() { struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + 1e-6 * t.tv_usec; }

// counters printed at exit (CHEMPS2_B200_VERBOSE=1) and readable by tests through the environment of the process
struct Counters {
   long long solves = 0, sigma_builds = 0, applies = 0, diags = 0, updates = 0, uploads = 0, registry_hits = 0;
   double t_solve = 0.0, t_update = 0.0, t_transfer = 0.0;
   ~Counters() {
      if (std::getenv("CHEMPS2_B200_VERBOSE"))
         std::fprintf(stderr, "chemps2_b200 drop-in: %lld Davidson solves (%lld sigma builds) %.2f s, %lld makeHeff, %lld fillHeffDiag, %lld operator updates %.2f s, "
                              "%lld operator-set uploads, %lld registry hits, host<->device operator traffic %.2f s\n",
                      solves, sigma_builds, t_solve, applies, diags, updates, t_update, uploads, registry_hits, t_transfer);
   }
} g_cnt;

// ---------------------------------------------------------------------------------------------------------------------------------
// one library context per Problem object (problem table + bookkeeper mirror)
struct Ctx {
   b2_ctx* h = nullptr;
   int L = 0, group = -1, N = -1, twoS = -1, irrep = -1;
   std::vector<int> orb_irrep;
   std::vector<double> mx;
};
std::map<const Problem*, Ctx> g_ctx;

int device_ordinal() { const char* e = std::getenv("CHEMPS2_B200_DEVICE"); return e ? std::atoi(e) : 0; }

// every device operator set the shim knows: which host tensors it mirrors
struct Entry {
   b2_opset* set = nullptr;
   const Problem* prob = nullptr;
   const TensorX* x = nullptr;      // the TensorX object of the table slot the set mirrors
   bool moving_right = true;
   int boundary = 0;
   uint64_t fingerprint = 0;        // content fingerprint of the host tensors at registration
   std::vector<int> dims;           // bookkeeper dimensions of the boundary at registration
   int64_t bytes = 0;
   long long last_use = 0;
};
long long g_tick = 0;
// device sets of DMRG objects that no longer exist (CASSCF builds one DMRG object per iteration) would pile up: least-recently-used
// sets beyond the budget are released (CHEMPS2_B200_HBM_GB, default 96); a released set is simply uploaded again when it is needed
double hbm_budget_bytes() { const char* e = std::getenv("CHEMPS2_B200_HBM_GB"); return (e ? std::atof(e) : 96.0) * 1073741824.0; }
std::map<std::pair<const void*, int>, Entry> g_sets;   // key: (the DMRG object's Xtensors table, slot)

uint64_t mix(uint64_t h, const void* p, size_t n) {   // FNV-1a over 8-byte words
   const uint64_t* w = static_cast<const uint64_t*>(p);
   for (size_t i = 0; i < n; i++) { h ^= w[i]; h *= 0x100000001B3ULL; }
   return h;
}
uint64_t tensor_fp(const Tensor* t, uint64_t h = 0xCBF29CE484222325ULL) {
   Tensor* tt = const_cast<Tensor*>(t);
   const int n = tt->gKappa2index(tt->gNKappa());
   h ^= (uint64_t)n; h *= 0x100000001B3ULL;
   return mix(h, tt->gStorage(), (size_t)n);
}

void drop_sets_of(const Problem* prob) {
   for (auto it = g_sets.begin(); it != g_sets.end();)
      if (it->second.prob == prob) { b2_opset_destroy(it->second.set); it = g_sets.erase(it); } else ++it;
}

std::vector<int> boundary_dims(const SyBookkeeper* bk, int b) {
   std::vector<int> d;
   for (int n = bk->gNmin(b); n <= bk->gNmax(b); n++)
      for (int ts = bk->gTwoSmin(b, n); ts <= bk->gTwoSmax(b, n); ts += 2)
         for (int ir = 0; ir < bk->getNumberOfIrreps(); ir++) d.push_back(bk->gCurrentDim(b, n, ts, ir));
   return d;
}

// context for (Problem, SyBookkeeper): created on first use; the folded integral table (Problem::gMxElement, which the reference's tests
// 9 and 12 rewrite through setMxElement) and every virtual dimension are re-synchronised on each entry — both are cheap next to a solve
b2_ctx* ctx_for(const Problem* prob, const SyBookkeeper* bk) {
   Ctx& c = g_ctx[prob];
   const int L = prob->gL();
   std::vector<int> irr(L);
   for (int i = 0; i < L; i++) irr[i] = prob->gIrrep(i);
   std::vector<double> mx((size_t)L * L * L * L);
   {
      size_t p = 0;   // mx[a + L*(b + L*(c + L*d))]
      for (int d = 0; d < L; d++) for (int cc = 0; cc < L; cc++) for (int b = 0; b < L; b++) for (int a = 0; a < L; a++) mx[p++] = prob->gMxElement(a, b, cc, d);
   }
   const bool same_shape = c.h && c.L == L && c.group == prob->gSy() && c.N == prob->gN() && c.twoS == prob->gTwoS() && c.irrep == prob->gIrrep() && c.orb_irrep == irr;
   if (!same_shape) {
      if (c.h) { drop_sets_of(prob); b2_ctx_destroy(c.h); c.h = nullptr; }
      B2(b2_ctx_create(device_ordinal(), &c.h));
      B2(b2_problem_set(c.h, L, prob->gSy(), prob->gN(), prob->gTwoS(), prob->gIrrep(), irr.data(), mx.data(), prob->gEconst()));
      B2(b2_bk_init(c.h, 1));
      c.L = L; c.group = prob->gSy(); c.N = prob->gN(); c.twoS = prob->gTwoS(); c.irrep = prob->gIrrep(); c.orb_irrep = irr; c.mx = mx;
   } else if (c.mx != mx) {
      B2(b2_problem_update_mx(c.h, mx.data()));
      c.mx = mx;
   }
   for (int b = 0; b <= L; b++)
      for (int n = bk->gNmin(b); n <= bk->gNmax(b); n++)
         for (int ts = bk->gTwoSmin(b, n); ts <= bk->gTwoSmax(b, n); ts += 2)
            for (int ir = 0; ir < bk->getNumberOfIrreps(); ir++) B2(b2_bk_set_dim(c.h, b, n, ts, ir, bk->gCurrentDim(b, n, ts, ir)));
   return c.h;
}

// the reference's operator tables of one slot (DMRG.h:211-226), as Heff / DMRG hand them around
struct Tables {
   TensorL*** L; TensorOperator**** A; TensorOperator**** B; TensorOperator**** C; TensorOperator**** D;
   TensorS0**** S0; TensorS1**** S1; TensorF0**** F0; TensorF1**** F1; TensorQ*** Q; TensorX** X;
};

// the host tensor behind operator (kind, i, j) of table slot t — index conventions of DMRG::allocateTensors (DMRGoperators.cpp:909-1145)
Tensor* host_tensor(const Tables& T, int t, bool mr, int kind, int i, int j) {
   const bool inside = (kind == B2_L || kind == B2_S0 || kind == B2_S1 || kind == B2_F0 || kind == B2_F1);
   if (kind == B2_X) return T.X[t];
   if (kind == B2_L) return mr ? T.L[t][t - i] : T.L[t][i - (t + 1)];
   if (kind == B2_Q) return mr ? T.Q[t][i - (t + 1)] : T.Q[t][t - i];
   int c2 = j - i, c3;
   if (inside) c3 = mr ? t - j : i - (t + 1);
   else c3 = mr ? i - (t + 1) : t - j;
   switch (kind) {
      case B2_S0: return T.S0[t][c2][c3];
      case B2_S1: return T.S1[t][c2][c3];
      case B2_F0: return T.F0[t][c2][c3];
      case B2_F1: return T.F1[t][c2][c3];
      case B2_A: return T.A[t][c2][c3];
      case B2_B: return T.B[t][c2][c3];
      case B2_C: return T.C[t][c2][c3];
      case B2_D: return T.D[t][c2][c3];
   }
   return NULL;
}

// fingerprint of a slot: the X tensor, the first L tensor and one two-operator tensor — enough to tell whether the host tensors still
// hold what the device set was computed from / downloaded into
uint64_t slot_fp(const Tables& T, int t, bool mr, int L) {
   uint64_t h = tensor_fp(T.X[t]);
   h = tensor_fp(T.L[t][0], h);
   h = tensor_fp(T.F0[t][0][0], h);
   const int n_out = mr ? L - 1 - t : t + 1;
   if (n_out > 0) h = tensor_fp(T.Q[t][0], h);
   return h;
}

void transfer_all(b2_opset* set, const Tables& T, int t, bool mr, bool to_device) {
   const double t0 = wall();
   const int n = b2_opset_count(set);
   for (int idx = 0; idx < n; idx++) {
      int kind, i, j; int64_t size;
      B2(b2_opset_info(set, idx, &kind, &i, &j, &size));
      Tensor* ht = host_tensor(T, t, mr, kind, i, j);
      if (!ht || (int64_t)ht->gKappa2index(ht->gNKappa()) != size) {
         std::fprintf(stderr, "chemps2_b200 drop-in: operator kind %d (%d,%d) of slot %d: host tensor %s\n", kind, i, j, t, ht ? "has another size" : "missing");
         std::abort();
      }
      if (to_device) B2(b2_opset_upload(set, idx, ht->gStorage())); else B2(b2_opset_download(set, idx, ht->gStorage()));
   }
   g_cnt.t_transfer += wall() - t0;
}

void register_set(const Tables& T, int t, bool mr, b2_opset* set, const Problem* prob, const SyBookkeeper* bk) {
   auto key = std::make_pair((const void*)T.X, t);
   auto it = g_sets.find(key);
   if (it != g_sets.end()) { b2_opset_destroy(it->second.set); g_sets.erase(it); }
   Entry e;
   e.set = set; e.prob = prob; e.x = T.X[t]; e.moving_right = mr; e.boundary = t + 1;
   e.fingerprint = slot_fp(T, t, mr, prob->gL()); e.dims = boundary_dims(bk, t + 1);
   e.bytes = 8 * b2_opset_arena_size(set); e.last_use = ++g_tick;
   g_sets[key] = e;
   double total = 0.0;
   for (auto& kv : g_sets) total += (double)kv.second.bytes;
   while (total > hbm_budget_bytes() && g_sets.size() > 3) {   // never below the three sets one sweep step touches
      auto victim = g_sets.end();
      for (auto i2 = g_sets.begin(); i2 != g_sets.end(); ++i2)
         if (i2->second.last_use + 3 <= g_tick && (victim == g_sets.end() || i2->second.last_use < victim->second.last_use)) victim = i2;
      if (victim == g_sets.end()) break;
      total -= (double)victim->second.bytes;
      b2_opset_destroy(victim->second.set);
      g_sets.erase(victim);
   }
}

// device operator set for table slot t: the registered one when it still mirrors the host tensors, else a fresh upload
b2_opset* opset_for(b2_ctx* ctx, const Tables& T, int t, bool mr, const Problem* prob, const SyBookkeeper* bk) {
   auto key = std::make_pair((const void*)T.X, t);
   auto it = g_sets.find(key);
   if (it != g_sets.end()) {
      const Entry& e = it->second;
      if (e.prob == prob && e.x == T.X[t] && e.moving_right == mr && e.dims == boundary_dims(bk, t + 1) && e.fingerprint == slot_fp(T, t, mr, prob->gL())) {
         g_cnt.registry_hits++;
         it->second.last_use = ++g_tick;
         return e.set;
      }
   }
   b2_opset* set = nullptr;
   B2(b2_opset_create(ctx, t + 1, mr ? 1 : 0, &set));
   transfer_all(set, T, t, mr, true);
   g_cnt.uploads++;
   register_set(T, t, mr, set, prob, bk);
   return set;
}

// sigma plan of the last call, kept while the same Sobject / operators are used (the reference's Davidson calls makeHeff in a loop)
struct HeffCache {
   b2_heff* h = nullptr;
   const Problem* prob = nullptr; const void* xtab = nullptr; int index = -1;
   b2_opset *left = nullptr, *right = nullptr;
   std::vector<int> dims;
   int n_lower = 0;
} g_heff;

b2_heff* heff_for(const SyBookkeeper* bk, const Problem* prob, const Sobject* denS, const Tables& T, int nLower, double** VeffTilde) {
   const int index = denS->gIndex(), L = prob->gL();
   b2_ctx* ctx = ctx_for(prob, bk);
   b2_opset* left = index > 0 ? opset_for(ctx, T, index - 1, true, prob, bk) : nullptr;
   b2_opset* right = index < L - 2 ? opset_for(ctx, T, index + 1, false, prob, bk) : nullptr;
   std::vector<int> dims;
   for (int b = index; b <= index + 2; b++) { std::vector<int> d = boundary_dims(bk, b); dims.insert(dims.end(), d.begin(), d.end()); }
   const bool reuse = g_heff.h && g_heff.prob == prob && g_heff.xtab == (const void*)T.X && g_heff.index == index && g_heff.left == left && g_heff.right == right &&
                      g_heff.dims == dims;
   if (!reuse) {
      b2_heff_destroy(g_heff.h);
      g_heff = HeffCache();
      B2(b2_heff_create(ctx, index, left, right, 1, 0, &g_heff.h));
      g_heff.prob = prob; g_heff.xtab = (const void*)T.X; g_heff.index = index; g_heff.left = left; g_heff.right = right; g_heff.dims = dims;
   }
   if (nLower > 0 || g_heff.n_lower > 0) {   // the level-shift projector of the excited-state calculations (Heff.h:70 nLower / VeffTilde)
      B2(b2_heff_set_excitations(g_heff.h, nLower, (const double* const*)VeffTilde));
      g_heff.n_lower = nLower;
   }
   return g_heff.h;
}

void forget_heff() { b2_heff_destroy(g_heff.h); g_heff = HeffCache(); }

}   // namespace

// =================================================================================================================================
// Heff (Heff.h:50-100)
double CheMPS2::Heff::SolveDAVIDSON(Sobject* denS, TensorL*** Ltensors, TensorOperator**** Atensors, TensorOperator**** Btensors, TensorOperator**** Ctensors,
                                    TensorOperator**** Dtensors, TensorS0**** S0tensors, TensorS1**** S1tensors, TensorF0**** F0tensors, TensorF1**** F1tensors,
                                    TensorQ*** Qtensors, TensorX** Xtensors, int nLower, double** VeffTilde) const {
   const double t0 = wall();
   const Tables T = {Ltensors, Atensors, Btensors, Ctensors, Dtensors, S0tensors, S1tensors, F0tensors, F1tensors, Qtensors, Xtensors};
   b2_heff* h = heff_for(denBK, Prob, denS, T, nLower, VeffTilde);
   if (b2_heff_veclength(h) != (int64_t)denS->gKappa2index(denS->gNKappa())) die("Sobject layout differs from the library's", -1);
   double eigenvalue = 0.0;
   int nmv = 0;
   // the device Davidson (CheMPS2::Davidson's algorithm, Options.h:70-72 constants) on the Sobject storage in the program convention
   B2(b2_heff_solve(h, denS->gStorage(), dvdson_rtol, &eigenvalue, &nmv));
   if (CheMPS2::HEFF_debugPrint) { std::cout << "   Stats: nIt(DAVIDSON) = " << nmv << std::endl; }
   forget_heff();   // the operators of this site pair are replaced right after the solve: release the plan's workspaces now
   g_cnt.solves++; g_cnt.sigma_builds += nmv; g_cnt.t_solve += wall() - t0;
   return eigenvalue;
}

void CheMPS2::Heff::makeHeff(double* memS, double* memHeff, const Sobject* denS, TensorL*** Ltensors, TensorOperator**** Atensors, TensorOperator**** Btensors,
                             TensorOperator**** Ctensors, TensorOperator**** Dtensors, TensorS0**** S0tensors, TensorS1**** S1tensors, TensorF0**** F0tensors,
                             TensorF1**** F1tensors, TensorQ*** Qtensors, TensorX** Xtensors, int nLower, double** VeffTilde) const {
   const Tables T = {Ltensors, Atensors, Btensors, Ctensors, Dtensors, S0tensors, S1tensors, F0tensors, F1tensors, Qtensors, Xtensors};
   b2_heff* h = heff_for(denBK, Prob, denS, T, nLower, VeffTilde);
   B2(b2_heff_apply(h, memS, memHeff));
   g_cnt.applies++;
}

void CheMPS2::Heff::fillHeffDiag(double* memHeffDiag, const Sobject* denS, TensorOperator**** Ctensors, TensorOperator**** Dtensors, TensorF0**** F0tensors,
                                 TensorF1**** F1tensors, TensorX** Xtensors, int nLower, double** VeffTilde) const {
   // the diagonal needs C, D, F0, F1 and X only, but the sigma plan it rides on mirrors whole table slots: the other tables are reached
   // through the registry (the sets were registered by the operator updates); without them this entry cannot be served
   auto lookup = [&](int t) -> const Entry* { auto it = g_sets.find(std::make_pair((const void*)Xtensors, t)); return it == g_sets.end() ? nullptr : &it->second; };
   const int index = denS->gIndex(), L = Prob->gL();
   if (g_heff.h && g_heff.prob == Prob && g_heff.xtab == (const void*)Xtensors && g_heff.index == index) {
      if (nLower > 0 || g_heff.n_lower > 0) { B2(b2_heff_set_excitations(g_heff.h, nLower, (const double* const*)VeffTilde)); g_heff.n_lower = nLower; }
      B2(b2_heff_diag(g_heff.h, memHeffDiag));
      g_cnt.diags++;
      return;
   }
   b2_ctx* ctx = ctx_for(Prob, denBK);
   const Entry* el = index > 0 ? lookup(index - 1) : nullptr;
   const Entry* er = index < L - 2 ? lookup(index + 1) : nullptr;
   if ((index > 0 && !el) || (index < L - 2 && !er)) die("Heff::fillHeffDiag before any sigma build / operator update of this site pair", -1);
   b2_heff_destroy(g_heff.h);
   g_heff = HeffCache();
   B2(b2_heff_create(ctx, index, el ? el->set : nullptr, er ? er->set : nullptr, 1, 0, &g_heff.h));
   g_heff.prob = Prob; g_heff.xtab = (const void*)Xtensors; g_heff.index = index; g_heff.left = el ? el->set : nullptr; g_heff.right = er ? er->set : nullptr;
   for (int b = index; b <= index + 2; b++) { std::vector<int> d = boundary_dims(denBK, b); g_heff.dims.insert(g_heff.dims.end(), d.begin(), d.end()); }
   if (nLower > 0) { B2(b2_heff_set_excitations(g_heff.h, nLower, (const double* const*)VeffTilde)); g_heff.n_lower = nLower; }
   B2(b2_heff_diag(g_heff.h, memHeffDiag));
   g_cnt.diags++;
}

// =================================================================================================================================
// DMRG::updateMovingRight / updateMovingLeft (DMRG.h: private; called by the updateMoving*Safe* wrappers of DMRGoperators.cpp:33-231)
namespace {

void run_update(CheMPS2::Problem* Prob, const SyBookkeeper* denBK, const Tables& T, int L, TensorT** MPS, int index, bool mr) {
   b2_ctx* ctx = ctx_for(Prob, denBK);
   forget_heff();
   const int t_new = index;                        // table slot that receives the new operators (boundary index + 1)
   const int t_old = mr ? index - 1 : index + 1;   // slot of the operators one site further out
   const int site = mr ? index : index + 1;        // the MPS tensor that was just optimised
   const bool have_old = mr ? (index > 0) : (index < L - 2);
   b2_opset* old_set = have_old ? opset_for(ctx, T, t_old, mr, Prob, denBK) : nullptr;
   b2_opset* fresh = nullptr;
   B2(b2_opset_create(ctx, t_new + 1, mr ? 1 : 0, &fresh));
   b2_update* u = nullptr;
   B2(b2_update_create(ctx, site, mr ? 1 : 0, old_set, fresh, &u));
   B2(b2_update_run(u, MPS[site]->gStorage()));
   b2_update_destroy(u);
   transfer_all(fresh, T, t_new, mr, false);       // host-mirror rule: the caller's tensors receive the new operators
   register_set(T, t_new, mr, fresh, Prob, denBK);
   g_cnt.updates++;
}

}   // namespace

void CheMPS2::DMRG::updateMovingRight(const int index) {
   const double t0 = wall();
   const Tables T = {Ltensors, Atensors, Btensors, Ctensors, Dtensors, S0tensors, S1tensors, F0tensors, F1tensors, Qtensors, Xtensors};
   run_update(Prob, denBK, T, L, MPS, index, true);
   if (Exc_activated) {   // overlaps with the lower states ride along on the host (TensorO, tiny): same calls as DMRGoperators.cpp:556-567
      for (int state = 0; state < nStates - 1; state++) {
         TensorO* o = Exc_Overlaps[state][index];
         if (index == 0) o->create(MPS[index], Exc_MPSs[state][index]);
         else o->update_ownmem(MPS[index], Exc_MPSs[state][index], Exc_Overlaps[state][index - 1]);
      }
   }
   const double dt = wall() - t0;
   timings[CHEMPS2_TIME_TENS_CALC] += dt;
   g_cnt.t_update += dt;
}

void CheMPS2::DMRG::updateMovingLeft(const int index) {
   const double t0 = wall();
   const Tables T = {Ltensors, Atensors, Btensors, Ctensors, Dtensors, S0tensors, S1tensors, F0tensors, F1tensors, Qtensors, Xtensors};
   run_update(Prob, denBK, T, L, MPS, index, false);
   if (Exc_activated) {   // DMRGoperators.cpp:889-900
      for (int state = 0; state < nStates - 1; state++) {
         TensorO* o = Exc_Overlaps[state][index];
         if (index == L - 2) o->create(MPS[index + 1], Exc_MPSs[state][index + 1]);
         else o->update_ownmem(MPS[index + 1], Exc_MPSs[state][index + 1], Exc_Overlaps[state][index + 1]);
      }
   }
   const double dt = wall() - t0;
   timings[CHEMPS2_TIME_TENS_CALC] += dt;
   g_cnt.t_update += dt;
}
