/* TEST INFRASTRUCTURE ONLY — never linked into or called by the product.
 *
 * Drives the UNMODIFIED CheMPS2 reference (oracle/_ref/libchemps2.so) through its own classes to
 *   (1) dump golden fixtures of the two-site sweep hot path: bookkeeper dims, MPS tensors, every renormalized
 *       operator at the two boundaries of a site pair, the vector going into and coming out of Heff::makeHeff,
 *       the Heff diagonal, and operator sets before/after updateMovingLeft/Right;
 *   (2) print per-sweep energies / discarded weights of a full DMRG::Solve() for end-to-end parity;
 *   (3) time Heff::makeHeff (the CPU baseline of bench.py).
 *
 * Private members of the reference classes are reached with the `#define private public` trick; this file
 * contains no reference source code.
 *
 * Fixture container (".b2fx"): records of { int32 name_len, name, int32 dtype (0=int32,1=float64), int64 count, data }.
 */
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>
#include <iostream>
#include <sstream>
#include <sys/time.h>
#include <unistd.h>
#include <omp.h>

#define private public
#define protected public
#include "Initialize.h"
#include "Hamiltonian.h"
#include "Problem.h"
#include "ConvergenceScheme.h"
#include "SyBookkeeper.h"
#include "TensorT.h"
#include "TensorOperator.h"
#include "TensorL.h"
#include "TensorX.h"
#include "TensorQ.h"
#include "TensorS0.h"
#include "TensorS1.h"
#include "TensorF0.h"
#include "TensorF1.h"
#include "Sobject.h"
#include "Heff.h"
#include "DMRG.h"
#include "TwoDM.h"
#include "Correlations.h"
#include "Wigner.h"
#undef private
#undef protected

using namespace CheMPS2;

struct Writer {
   FILE * f;
   explicit Writer(const std::string & path){ f = fopen(path.c_str(), "wb"); if (!f){ perror(path.c_str()); exit(2); } }
   ~Writer(){ fclose(f); }
   void rec(const std::string & name, int dtype, long long n, const void * data){
      int len = (int) name.size();
      fwrite(&len, 4, 1, f); fwrite(name.data(), 1, len, f); fwrite(&dtype, 4, 1, f); fwrite(&n, 8, 1, f);
      fwrite(data, dtype == 0 ? 4 : 8, n, f);
   }
   void ints(const std::string & name, const std::vector<int> & v){ rec(name, 0, v.size(), v.data()); }
   void dbls(const std::string & name, const std::vector<double> & v){ rec(name, 1, v.size(), v.data()); }
   void dbls(const std::string & name, const double * p, long long n){ rec(name, 1, n, p); }
};

enum { K_L = 0, K_S0, K_S1, K_F0, K_F1, K_A, K_B, K_C, K_D, K_Q, K_X, K_G, K_Y, K_Z, K_K, K_M };

static void push_op(std::vector<int> & meta, std::vector<double> & data, int kind, int i, int j, Tensor * t){
   if (t == NULL) return;
   const int size = t->gKappa2index(t->gNKappa());
   meta.push_back(kind); meta.push_back(i); meta.push_back(j); meta.push_back(size);
   data.insert(data.end(), t->gStorage(), t->gStorage() + size);
}

/* All operators in table slot t (boundary t+1). Index conventions: DMRGoperators.cpp:909-1140. */
static void dump_ops(Writer & w, const std::string & prefix, DMRG & d, int t, bool moving_right){
   std::vector<int> meta; std::vector<double> data;
   const int L = d.L;
   if (moving_right){
      for (int k = 0; k <= t; k++) push_op(meta, data, K_L, t - k, t - k, d.Ltensors[t][k]);
      for (int c2 = 0; c2 <= t; c2++) for (int c3 = 0; c3 <= t - c2; c3++){
         const int j = t - c3, i = j - c2;
         push_op(meta, data, K_S0, i, j, d.S0tensors[t][c2][c3]);
         if (c2 > 0) push_op(meta, data, K_S1, i, j, d.S1tensors[t][c2][c3]);
         push_op(meta, data, K_F0, i, j, d.F0tensors[t][c2][c3]);
         push_op(meta, data, K_F1, i, j, d.F1tensors[t][c2][c3]);
      }
      for (int c2 = 0; c2 < L - 1 - t; c2++) for (int c3 = 0; c3 < L - 1 - t - c2; c3++){
         const int i = t + 1 + c3, j = i + c2;
         push_op(meta, data, K_A, i, j, d.Atensors[t][c2][c3]);
         if (c2 > 0) push_op(meta, data, K_B, i, j, d.Btensors[t][c2][c3]);
         push_op(meta, data, K_C, i, j, d.Ctensors[t][c2][c3]);
         push_op(meta, data, K_D, i, j, d.Dtensors[t][c2][c3]);
      }
      for (int c2 = 0; c2 < L - 1 - t; c2++) push_op(meta, data, K_Q, t + 1 + c2, t + 1 + c2, d.Qtensors[t][c2]);
      push_op(meta, data, K_X, -1, -1, d.Xtensors[t]);
   } else {
      for (int k = 0; k < L - 1 - t; k++) push_op(meta, data, K_L, t + 1 + k, t + 1 + k, d.Ltensors[t][k]);
      for (int c2 = 0; c2 < L - 1 - t; c2++) for (int c3 = 0; c3 < L - 1 - t - c2; c3++){
         const int i = t + 1 + c3, j = i + c2;
         push_op(meta, data, K_S0, i, j, d.S0tensors[t][c2][c3]);
         if (c2 > 0) push_op(meta, data, K_S1, i, j, d.S1tensors[t][c2][c3]);
         push_op(meta, data, K_F0, i, j, d.F0tensors[t][c2][c3]);
         push_op(meta, data, K_F1, i, j, d.F1tensors[t][c2][c3]);
      }
      for (int c2 = 0; c2 <= t; c2++) for (int c3 = 0; c3 <= t - c2; c3++){
         const int j = t - c3, i = j - c2;
         push_op(meta, data, K_A, i, j, d.Atensors[t][c2][c3]);
         if (c2 > 0) push_op(meta, data, K_B, i, j, d.Btensors[t][c2][c3]);
         push_op(meta, data, K_C, i, j, d.Ctensors[t][c2][c3]);
         push_op(meta, data, K_D, i, j, d.Dtensors[t][c2][c3]);
      }
      for (int c2 = 0; c2 <= t; c2++) push_op(meta, data, K_Q, t - c2, t - c2, d.Qtensors[t][c2]);
      push_op(meta, data, K_X, -1, -1, d.Xtensors[t]);
   }
   std::vector<int> hdr; hdr.push_back(t + 1); hdr.push_back(moving_right ? 1 : 0);
   w.ints(prefix + "/hdr", hdr);   /* boundary, moving_right */
   w.ints(prefix + "/meta", meta);
   w.dbls(prefix + "/data", data);
}

static void dump_bk(Writer & w, const std::string & prefix, const SyBookkeeper * bk){
   std::vector<int> rows;   /* boundary, N, twoS, irrep, curdim, fcidim */
   for (int b = 0; b <= bk->gL(); b++)
      for (int N = bk->gNmin(b); N <= bk->gNmax(b); N++)
         for (int ts = bk->gTwoSmin(b, N); ts <= bk->gTwoSmax(b, N); ts += 2)
            for (int ir = 0; ir < bk->getNumberOfIrreps(); ir++){
               rows.push_back(b); rows.push_back(N); rows.push_back(ts); rows.push_back(ir);
               rows.push_back(bk->gCurrentDim(b, N, ts, ir)); rows.push_back(bk->gFCIdim(b, N, ts, ir));
            }
   w.ints(prefix, rows);
}

static void dump_mps(Writer & w, const std::string & prefix, DMRG & d){
   for (int s = 0; s < d.L; s++){
      std::ostringstream nm; nm << prefix << "/" << s;
      w.dbls(nm.str(), d.MPS[s]->gStorage(), d.MPS[s]->gKappa2index(d.MPS[s]->gNKappa()));
   }
}

static void dump_problem(Writer & w, Problem * prob, Hamiltonian * ham, int group){
   const int L = prob->gL();
   std::vector<int> hdr; hdr.push_back(L); hdr.push_back(group); hdr.push_back(prob->gN()); hdr.push_back(prob->gTwoS()); hdr.push_back(prob->gIrrep());
   w.ints("problem/hdr", hdr);
   std::vector<int> irr; for (int i = 0; i < L; i++) irr.push_back(prob->gIrrep(i));
   w.ints("problem/orb_irrep", irr);
   std::vector<double> mx((size_t) L * L * L * L), tm((size_t) L * L), vm((size_t) L * L * L * L);
   for (int a = 0; a < L; a++) for (int b = 0; b < L; b++){
      tm[a + L * b] = ham->getTmat(prob->bReorder ? prob->f2[a] : a, prob->bReorder ? prob->f2[b] : b);
      for (int c = 0; c < L; c++) for (int e = 0; e < L; e++){
         mx[a + L * (b + L * (c + (size_t) L * e))] = prob->gMxElement(a, b, c, e);
         vm[a + L * (b + L * (c + (size_t) L * e))] = ham->getVmat(prob->bReorder ? prob->f2[a] : a, prob->bReorder ? prob->f2[b] : b,
                                                                    prob->bReorder ? prob->f2[c] : c, prob->bReorder ? prob->f2[e] : e);
      }
   }
   w.dbls("problem/mx", mx); w.dbls("problem/tmat", tm); w.dbls("problem/vmat", vm);
   std::vector<double> ec; ec.push_back(prob->gEconst()); w.dbls("problem/econst", ec);
   bool direct = false;   /* table not derivable from (T, V): written with setMxElement */
   for (int a = 0; a < L && !direct; a++) for (int b = 0; b < L && !direct; b++) for (int c = 0; c < L && !direct; c++) for (int e = 0; e < L; e++){
      const double fold = vm[a + L * (b + L * (c + (size_t) L * e))] + ((a == c) ? tm[b + L * e] : 0.0) / (prob->gN() - 1.0) + ((b == e) ? tm[a + L * c] : 0.0) / (prob->gN() - 1.0);
      if (fabs(fold - mx[a + L * (b + L * (c + (size_t) L * e))]) > 1e-12){ direct = true; break; }
   }
   std::vector<int> dm; dm.push_back(direct ? 1 : 0); w.ints("problem/direct_mx", dm);
}

/* sigma case at site `index`: S (symmetric convention), H*S, diag(H) straight from Heff::makeHeff / fillHeffDiag */
static void dump_sigma(Writer & w, const std::string & prefix, DMRG & d, int index){
   Sobject S(index, d.denBK);
   S.Join(d.MPS[index], d.MPS[index + 1]);
   const int n = S.gKappa2index(S.gNKappa());
   w.dbls(prefix + "/joined", S.gStorage(), n);   /* program convention, output of Join */
   S.prog2symm();
   std::vector<double> out(n), diag(n);
   Heff solver(d.denBK, d.Prob, 1e-5);
   solver.makeHeff(S.gStorage(), out.data(), &S, d.Ltensors, d.Atensors, d.Btensors, d.Ctensors, d.Dtensors, d.S0tensors, d.S1tensors,
                   d.F0tensors, d.F1tensors, d.Qtensors, d.Xtensors, 0, NULL);
   solver.fillHeffDiag(diag.data(), &S, d.Ctensors, d.Dtensors, d.F0tensors, d.F1tensors, d.Xtensors, 0, NULL);
   std::vector<int> hdr; hdr.push_back(index); hdr.push_back(n); hdr.push_back(S.gNKappa());
   w.ints(prefix + "/hdr", hdr);
   w.dbls(prefix + "/vec_in", S.gStorage(), n);
   w.dbls(prefix + "/vec_out", out);
   w.dbls(prefix + "/diag", diag);
   /* a second, random input so that parity does not hinge on the structure of the joined state */
   std::vector<double> rin(n), rout(n);
   unsigned int st = 12345u + index;
   for (int i = 0; i < n; i++){ st = st * 1664525u + 1013904223u; rin[i] = ((st >> 8) & 0xFFFF) / 65536.0 - 0.5; }
   solver.makeHeff(rin.data(), rout.data(), &S, d.Ltensors, d.Atensors, d.Btensors, d.Ctensors, d.Dtensors, d.S0tensors, d.S1tensors,
                   d.F0tensors, d.F1tensors, d.Qtensors, d.Xtensors, 0, NULL);
   w.dbls(prefix + "/rnd_in", rin);
   w.dbls(prefix + "/rnd_out", rout);
   /* excited-state level-shift projector (Heff::addDiagramExcitations, HeffDiagrams1.cpp:65-85; addDiagonalExcitations,
      HeffDiagonal.cpp:621-640): two random "VeffTilde" vectors */
   {
      std::vector<double> v0(n), v1(n), eout(n), ediag(n);
      for (int i = 0; i < n; i++){ st = st * 1664525u + 1013904223u; v0[i] = ((st >> 8) & 0xFFFF) / 65536.0 - 0.5; }
      for (int i = 0; i < n; i++){ st = st * 1664525u + 1013904223u; v1[i] = ((st >> 8) & 0xFFFF) / 65536.0 - 0.5; }
      double * vt[2] = { v0.data(), v1.data() };
      solver.makeHeff(rin.data(), eout.data(), &S, d.Ltensors, d.Atensors, d.Btensors, d.Ctensors, d.Dtensors, d.S0tensors, d.S1tensors,
                      d.F0tensors, d.F1tensors, d.Qtensors, d.Xtensors, 2, vt);
      solver.fillHeffDiag(ediag.data(), &S, d.Ctensors, d.Dtensors, d.F0tensors, d.F1tensors, d.Xtensors, 2, vt);
      w.dbls(prefix + "/exc_v0", v0); w.dbls(prefix + "/exc_v1", v1);
      w.dbls(prefix + "/exc_out", eout); w.dbls(prefix + "/exc_diag", ediag);
   }
   if (index > 0) dump_ops(w, prefix + "/left", d, index - 1, true);
   if (index < d.L - 2) dump_ops(w, prefix + "/right", d, index + 1, false);
}

/* G/Y/Z/K/M tensors of the two-orbital correlation functions along the chain: DMRG::update_correlations_tensors
   (DMRGoperators3RDM.cpp:415-479) called for siteindex = 1 .. L-1 on the current MPS, every table dumped */
static void dump_correlation_tensors(Writer & w, const std::string & prefix, DMRG & d){
   const int L = d.L;
   dump_bk(w, prefix + "/bk", d.denBK);
   dump_mps(w, prefix + "/mps", d);
   d.Gtensors = new TensorGYZ*[L - 1]; d.Ytensors = new TensorGYZ*[L - 1]; d.Ztensors = new TensorGYZ*[L - 1];
   d.Ktensors = new TensorKM*[L - 1];  d.Mtensors = new TensorKM*[L - 1];
   for (int siteindex = 1; siteindex < L; siteindex++){
      d.update_correlations_tensors(siteindex);
      std::vector<int> meta; std::vector<double> data;
      for (int prev = 0; prev < siteindex; prev++){
         push_op(meta, data, K_G, prev, prev, d.Gtensors[prev]);
         push_op(meta, data, K_Y, prev, prev, d.Ytensors[prev]);
         push_op(meta, data, K_Z, prev, prev, d.Ztensors[prev]);
         push_op(meta, data, K_K, prev, prev, d.Ktensors[prev]);
         push_op(meta, data, K_M, prev, prev, d.Mtensors[prev]);
      }
      std::ostringstream nm; nm << prefix << "/b" << siteindex;
      std::vector<int> hdr; hdr.push_back(siteindex); hdr.push_back(1);
      w.ints(nm.str() + "/hdr", hdr); w.ints(nm.str() + "/meta", meta); w.dbls(nm.str() + "/data", data);
   }
   for (int prev = 0; prev < L - 1; prev++){ delete d.Gtensors[prev]; delete d.Ytensors[prev]; delete d.Ztensors[prev]; delete d.Ktensors[prev]; delete d.Mtensors[prev]; }
   delete [] d.Gtensors; delete [] d.Ytensors; delete [] d.Ztensors; delete [] d.Ktensors; delete [] d.Mtensors;
   d.Gtensors = NULL; d.Ytensors = NULL; d.Ztensors = NULL; d.Ktensors = NULL; d.Mtensors = NULL;
}

struct Setup {
   Hamiltonian * ham; Problem * prob; int group;
   int pairingL; double pairing_g, pairing_power;   /* --pairing: the reduced BCS model of the reference's tests/test12.cpp.in */
   int hub2d; double hub2d_U, hub2d_T; bool momentum; /* --hubbard2d Llinear U T [--momentum]: the square Hubbard model with PBC of tests/test9.cpp.in */
   Setup() : ham(NULL), prob(NULL), group(0), pairingL(0), pairing_g(0.0), pairing_power(0.0), hub2d(0), hub2d_U(0.0), hub2d_T(0.0), momentum(false) {}
};

/* tests/test9.cpp.in:83-113: the momentum-space form of the same model, written directly into the folded table (plane-wave orbitals:
   the table has only 4-fold permutation symmetry); (k1 + k2 = k3 + k4 mod Llinear) per direction, dispersion 2T(cos kx + cos ky) */
static void apply_momentum_hubbard(const Setup & s, DMRG & d){
   if (s.hub2d <= 0 || !s.momentum) return;
   const int n = s.hub2d, L = n * n, N = s.prob->gN();
   std::vector<double> disp(L);
   for (int o = 0; o < L; o++) disp[o] = 2 * s.hub2d_T * (cos(2 * M_PI * (o % n) / n) + cos(2 * M_PI * (o / n) / n));
   for (int o1 = 0; o1 < L; o1++) for (int o2 = 0; o2 < L; o2++) for (int o3 = 0; o3 < L; o3++) for (int o4 = 0; o4 < L; o4++){
      const bool kx = ((o1 % n) + (o2 % n)) % n == ((o3 % n) + (o4 % n)) % n;
      const bool ky = ((o1 / n) + (o2 / n)) % n == ((o3 / n) + (o4 / n)) % n;
      double v = (kx && ky) ? s.hub2d_U / L : 0.0;
      if (o1 == o3 && o2 == o4) v += (disp[o1] + disp[o2]) / (N - 1);
      s.prob->setMxElement(o1, o2, o3, o4, v);
   }
   d.PreSolve();
}

/* tests/test12.cpp.in:57-79: the folded table is written DIRECTLY with Problem::setMxElement after the DMRG object exists (the table
   is not 8-fold symmetric, so it cannot go through Hamiltonian::setVmat), then PreSolve rebuilds the operators */
static void apply_pairing_model(const Setup & s, DMRG & d){
   if (s.pairingL <= 0) return;
   const int L = s.pairingL; const int N = s.prob->gN();
   for (int orb1 = 0; orb1 < L; orb1++){
      for (int orb2 = 0; orb2 < L; orb2++){
         const double e1 = -0.5 * (L - 1) + orb1, e2 = -0.5 * (L - 1) + orb2;   /* eps = -3.5 ... 3.5 for L = 8 */
         const double eri = s.pairing_g * pow(fabs(e1 * e2), s.pairing_power);
         const double oei = (e1 + e2) / (N - 1);
         if (orb1 == orb2){ s.prob->setMxElement(orb1, orb1, orb2, orb2, eri + oei); }
         else { s.prob->setMxElement(orb1, orb1, orb2, orb2, eri); s.prob->setMxElement(orb1, orb2, orb1, orb2, oei); }
      }
   }
   d.PreSolve();
}

static Setup make_setup(int argc, char ** argv){
   Setup s; std::string fcidump, problem; int twoS = 0, N = 0, irrep = 0, hubL = 0; double hubU = 0.0; bool reorder = false;
   for (int i = 2; i < argc; i++){
      std::string a = argv[i];
      if (a == "--fcidump") fcidump = argv[++i];
      else if (a == "--group") s.group = atoi(argv[++i]);
      else if (a == "--twoS") twoS = atoi(argv[++i]);
      else if (a == "--N") N = atoi(argv[++i]);
      else if (a == "--irrep") irrep = atoi(argv[++i]);
      else if (a == "--hubbard"){ hubL = atoi(argv[++i]); hubU = atof(argv[++i]); }
      else if (a == "--reorder") reorder = true;
      else if (a == "--problem") problem = argv[++i];
      else if (a == "--pairing"){ s.pairingL = atoi(argv[++i]); s.pairing_g = atof(argv[++i]); s.pairing_power = atof(argv[++i]); }
      else if (a == "--hubbard2d"){ s.hub2d = atoi(argv[++i]); s.hub2d_U = atof(argv[++i]); s.hub2d_T = atof(argv[++i]); }
      else if (a == "--momentum") s.momentum = true;
   }
   if (s.hub2d > 0){   /* site basis through the Hamiltonian class (8-fold symmetric there); the momentum form follows in apply_momentum_hubbard */
      const int n = s.hub2d, L2 = n * n;
      std::vector<int> irr(L2, 0);
      s.group = 0;
      s.ham = new Hamiltonian(L2, 0, irr.data());
      if (!s.momentum){
         for (int c = 0; c < L2; c++) s.ham->setVmat(c, c, c, c, s.hub2d_U);
         for (int ix = 0; ix < n; ix++) for (int iy = 0; iy < n; iy++){
            s.ham->setTmat(ix + n * iy, ((ix + 1) % n) + n * iy, s.hub2d_T);
            s.ham->setTmat(ix + n * iy, ix + n * ((iy + 1) % n), s.hub2d_T);
         }
      }
      s.prob = new Problem(s.ham, twoS, N, irrep);
      return s;
   }
   if (s.pairingL > 0){   /* all-zero Hamiltonian of the right shape; the matrix elements follow in apply_pairing_model */
      std::vector<int> irr(s.pairingL, 0);
      s.group = 0;
      s.ham = new Hamiltonian(s.pairingL, 0, irr.data());
      s.prob = new Problem(s.ham, twoS, N, irrep);
      return s;
   }
   if (!problem.empty()){   /* binary problem file of chemps2_b200/workloads.py (write_problem_file): synthetic / model Hamiltonians */
      FILE * f = fopen(problem.c_str(), "rb");
      if (!f){ perror(problem.c_str()); exit(2); }
      int hdr[5]; if (fread(hdr, 4, 5, f) != 5) exit(2);
      const int L = hdr[0]; s.group = hdr[1]; N = hdr[2]; twoS = hdr[3]; irrep = hdr[4];
      std::vector<int> irr(L); double econst = 0.0;
      std::vector<double> tm((size_t) L * L), vm((size_t) L * L * L * L);
      if (fread(irr.data(), 4, L, f) != (size_t) L || fread(&econst, 8, 1, f) != 1 || fread(tm.data(), 8, tm.size(), f) != tm.size() || fread(vm.data(), 8, vm.size(), f) != vm.size()) exit(2);
      fclose(f);
      s.ham = new Hamiltonian(L, s.group, irr.data());
      s.ham->setEconst(econst);
      for (int i = 0; i < L; i++) for (int j = i; j < L; j++) if (Irreps::directProd(irr[i], irr[j]) == 0) s.ham->setTmat(i, j, tm[i + (size_t) L * j]);
      for (int i = 0; i < L; i++) for (int j = 0; j < L; j++) for (int k = 0; k < L; k++) for (int l = 0; l < L; l++)
         if (Irreps::directProd(Irreps::directProd(irr[i], irr[j]), Irreps::directProd(irr[k], irr[l])) == 0) s.ham->setVmat(i, j, k, l, vm[i + L * (j + L * (k + (size_t) L * l))]);
      s.prob = new Problem(s.ham, twoS, N, irrep);
      return s;
   }
   if (hubL > 0){   /* 1-D Hubbard chain, open ends, C1 (pattern of the reference's tests/test4) */
      std::vector<int> irr(hubL, 0);
      s.group = 0;
      s.ham = new Hamiltonian(hubL, 0, irr.data());
      for (int i = 0; i < hubL; i++) for (int j = 0; j < hubL; j++){
         s.ham->setTmat(i, j, 0.0);
         for (int k = 0; k < hubL; k++) for (int l = 0; l < hubL; l++) s.ham->setVmat(i, j, k, l, 0.0);
      }
      s.ham->setEconst(0.0);
      for (int i = 0; i < hubL - 1; i++) s.ham->setTmat(i, i + 1, -1.0);
      for (int i = 0; i < hubL; i++) s.ham->setVmat(i, i, i, i, hubU);
   } else {
      s.ham = new Hamiltonian(fcidump, s.group);
   }
   s.prob = new Problem(s.ham, twoS, N, irrep);
   if (reorder && s.group == 7) s.prob->SetupReorderD2h();
   return s;
}

static int argi(int argc, char ** argv, const char * key, int def){ for (int i = 2; i < argc - 1; i++) if (!strcmp(argv[i], key)) return atoi(argv[i + 1]); return def; }
static double argd(int argc, char ** argv, const char * key, double def){ for (int i = 2; i < argc - 1; i++) if (!strcmp(argv[i], key)) return atof(argv[i + 1]); return def; }
static const char * args(int argc, char ** argv, const char * key, const char * def){ for (int i = 2; i < argc - 1; i++) if (!strcmp(argv[i], key)) return argv[i + 1]; return def; }

static double now(){ struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + 1e-6 * t.tv_usec; }


/* ------------------------------------------------------------------------------------------------------------------
 * "synth" mode: the reference's Heff::makeHeff on a synthetic workload WITHOUT running a DMRG calculation first.
 * The operator tables of the two boundaries next to the site pair are allocated with the reference's own constructors
 * (same index conventions as DMRG::allocateTensors, DMRGoperators.cpp:909-1145) and filled with a deterministic hash
 * of (seed, kind, site_i, site_j, element) — the same fill chemps2_b200's b2_opset_fill_hash produces — so the GPU
 * result can be compared with the reference's at the full benchmark size, and the reference can be timed on it. */
static inline double hash_value(unsigned long long seed, unsigned long long key, unsigned long long e){
   unsigned long long z = seed + 0x9E3779B97F4A7C15ULL * (key + 1) + 0xD1B54A32D192ED03ULL * (e + 1);
   z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
   z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
   z = z ^ (z >> 31);
   return (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
}
static inline unsigned long long op_key(int side, int kind, int i, int j){
   return ((unsigned long long) side << 60) | ((unsigned long long) kind << 40) | ((unsigned long long)(i + 1) << 20) | (unsigned long long)(j + 1);
}
static void fill_tensor(Tensor * t, unsigned long long seed, unsigned long long key, double amp){
   const long long n = t->gKappa2index(t->gNKappa());
   double * p = t->gStorage();
   #pragma omp parallel for schedule(static)
   for (long long e = 0; e < n; e++) p[e] = amp * hash_value(seed, key, e);
}

struct Tables {
   TensorL *** Lt; TensorF0 **** F0; TensorF1 **** F1; TensorS0 **** S0; TensorS1 **** S1;
   TensorOperator **** A; TensorOperator **** B; TensorOperator **** C; TensorOperator **** Dd; TensorQ *** Q; TensorX ** X;
   explicit Tables(int L){
      Lt = new TensorL ** [L]; F0 = new TensorF0 *** [L]; F1 = new TensorF1 *** [L]; S0 = new TensorS0 *** [L]; S1 = new TensorS1 *** [L];
      A = new TensorOperator *** [L]; B = new TensorOperator *** [L]; C = new TensorOperator *** [L]; Dd = new TensorOperator *** [L];
      Q = new TensorQ ** [L]; X = new TensorX * [L];
      for (int i = 0; i < L; i++){ Lt[i] = NULL; F0[i] = NULL; F1[i] = NULL; S0[i] = NULL; S1[i] = NULL; A[i] = NULL; B[i] = NULL; C[i] = NULL; Dd[i] = NULL; Q[i] = NULL; X[i] = NULL; }
   }
};

/* table slot t = boundary t+1; side = 1 (moving right, left of the site pair) or 2 (moving left, right of the pair) */
static void alloc_fill_side(Tables & T, SyBookkeeper * bk, Problem * prob, int t, bool mr, unsigned long long seed, double amp){
   const int L = bk->gL(); const int side = mr ? 1 : 2; const int b = t + 1;
   const int n_in  = mr ? t + 1 : L - 1 - t;   /* sites inside the renormalized block  */
   const int n_out = mr ? L - 1 - t : t + 1;   /* sites outside (complementary operators) */
   T.Lt[t] = new TensorL * [n_in];
   for (int k = 0; k < n_in; k++){
      const int s = mr ? t - k : t + 1 + k;
      T.Lt[t][k] = new TensorL(b, bk->gIrrep(s), mr, bk, bk); fill_tensor(T.Lt[t][k], seed, op_key(side, K_L, s, s), amp);
   }
   T.F0[t] = new TensorF0 ** [n_in]; T.F1[t] = new TensorF1 ** [n_in]; T.S0[t] = new TensorS0 ** [n_in]; T.S1[t] = new TensorS1 ** [n_in];
   for (int c2 = 0; c2 < n_in; c2++){
      T.F0[t][c2] = new TensorF0 * [n_in - c2]; T.F1[t][c2] = new TensorF1 * [n_in - c2]; T.S0[t][c2] = new TensorS0 * [n_in - c2];
      T.S1[t][c2] = (c2 > 0) ? new TensorS1 * [n_in - c2] : NULL;
      for (int c3 = 0; c3 < n_in - c2; c3++){
         int i, j; if (mr){ j = t - c3; i = j - c2; } else { i = t + 1 + c3; j = i + c2; }
         const int I = Irreps::directProd(bk->gIrrep(i), bk->gIrrep(j));
         T.F0[t][c2][c3] = new TensorF0(b, I, mr, bk); fill_tensor(T.F0[t][c2][c3], seed, op_key(side, K_F0, i, j), amp);
         T.F1[t][c2][c3] = new TensorF1(b, I, mr, bk); fill_tensor(T.F1[t][c2][c3], seed, op_key(side, K_F1, i, j), amp);
         T.S0[t][c2][c3] = new TensorS0(b, I, mr, bk); fill_tensor(T.S0[t][c2][c3], seed, op_key(side, K_S0, i, j), amp);
         if (c2 > 0){ T.S1[t][c2][c3] = new TensorS1(b, I, mr, bk); fill_tensor(T.S1[t][c2][c3], seed, op_key(side, K_S1, i, j), amp); }
      }
   }
   T.A[t] = new TensorOperator ** [n_out]; T.B[t] = new TensorOperator ** [n_out]; T.C[t] = new TensorOperator ** [n_out]; T.Dd[t] = new TensorOperator ** [n_out];
   for (int c2 = 0; c2 < n_out; c2++){
      T.A[t][c2] = new TensorOperator * [n_out - c2]; T.B[t][c2] = (c2 > 0) ? new TensorOperator * [n_out - c2] : NULL;
      T.C[t][c2] = new TensorOperator * [n_out - c2]; T.Dd[t][c2] = new TensorOperator * [n_out - c2];
      for (int c3 = 0; c3 < n_out - c2; c3++){
         int i, j; if (mr){ i = t + 1 + c3; j = i + c2; } else { j = t - c3; i = j - c2; }
         const int I = Irreps::directProd(bk->gIrrep(i), bk->gIrrep(j));
         T.A[t][c2][c3] = new TensorOperator(b, 0, 2, I, mr, true, false, bk, bk); fill_tensor(T.A[t][c2][c3], seed, op_key(side, K_A, i, j), amp);
         if (c2 > 0){ T.B[t][c2][c3] = new TensorOperator(b, 2, 2, I, mr, true, false, bk, bk); fill_tensor(T.B[t][c2][c3], seed, op_key(side, K_B, i, j), amp); }
         T.C[t][c2][c3] = new TensorOperator(b, 0, 0, I, mr, true, false, bk, bk); fill_tensor(T.C[t][c2][c3], seed, op_key(side, K_C, i, j), amp);
         T.Dd[t][c2][c3] = new TensorOperator(b, 2, 0, I, mr, mr, false, bk, bk); fill_tensor(T.Dd[t][c2][c3], seed, op_key(side, K_D, i, j), amp);
      }
   }
   T.Q[t] = new TensorQ * [n_out];
   for (int c2 = 0; c2 < n_out; c2++){
      const int s = mr ? t + 1 + c2 : t - c2;
      T.Q[t][c2] = new TensorQ(b, bk->gIrrep(s), mr, bk, prob, s); fill_tensor(T.Q[t][c2], seed, op_key(side, K_Q, s, s), amp);
   }
   T.X[t] = new TensorX(b, mr, bk, prob); fill_tensor(T.X[t], seed, op_key(side, K_X, -1, -1), amp);
}

static int run_synth(int argc, char ** argv){
   const char * pfile = args(argc, argv, "--problem", NULL);
   if (!pfile){ fprintf(stderr, "synth: --problem file needed\n"); return 1; }
   FILE * f = fopen(pfile, "rb"); if (!f){ perror(pfile); return 2; }
   int hdr[5]; if (fread(hdr, 4, 5, f) != 5) return 2;
   const int L = hdr[0], group = hdr[1], N = hdr[2], twoS = hdr[3], irrep = hdr[4];
   std::vector<int> irr(L); std::vector<double> tm((size_t) L * L), vm((size_t) L * L * L * L); double econst = 0.0;
   if (fread(irr.data(), 4, L, f) != (size_t) L || fread(&econst, 8, 1, f) != 1 || fread(tm.data(), 8, tm.size(), f) != tm.size() || fread(vm.data(), 8, vm.size(), f) != vm.size()) return 2;
   fclose(f);
   Hamiltonian ham(L, group, irr.data());
   ham.setEconst(econst);
   for (int a = 0; a < L; a++) for (int b = a; b < L; b++) if (irr[a] == irr[b]) ham.setTmat(a, b, tm[a + L * b]);
   for (int a = 0; a < L; a++) for (int b = 0; b < L; b++) for (int c = 0; c < L; c++) for (int e = 0; e < L; e++)
      if (Irreps::directProd(Irreps::directProd(irr[a], irr[b]), Irreps::directProd(irr[c], irr[e])) == 0) ham.setVmat(a, b, c, e, vm[a + L * (b + L * (c + (size_t) L * e))]);
   Problem prob(&ham, twoS, N, irrep);
   prob.construct_mxelem();
   const int D = argi(argc, argv, "--D", 100);
   const int site = argi(argc, argv, "--site", L / 2 - 1);
   const int reps = argi(argc, argv, "--reps", 1);
   const unsigned long long seed = (unsigned long long) argi(argc, argv, "--seed", 1);
   const double amp = argd(argc, argv, "--amp", 1.0);
   SyBookkeeper bk(&prob, D);
   const char * dfile = args(argc, argv, "--dims", NULL);
   if (dfile){   /* rows of int32: boundary N twoS irrep dim */
      FILE * g = fopen(dfile, "rb"); if (!g){ perror(dfile); return 2; }
      int row[5]; while (fread(row, 4, 5, g) == 5) bk.SetDim(row[0], row[1], row[2], row[3], row[4]);
      fclose(g);
   }
   const double t_alloc = now();
   Tables T(L);
   if (site > 0) alloc_fill_side(T, &bk, &prob, site - 1, true, seed, amp);
   if (site < L - 2) alloc_fill_side(T, &bk, &prob, site + 1, false, seed, amp);
   Sobject S(site, &bk);
   const long long n = S.gKappa2index(S.gNKappa());
   for (long long e = 0; e < n; e++) S.gStorage()[e] = hash_value(seed, op_key(3, 0, -1, -1), e);
   std::vector<double> out(n), diag(n);
   Heff solver(&bk, &prob, 1e-5);
   const double t_setup = now() - t_alloc;
   double best = 1e99, tot = 0.0;
   for (int r = 0; r < reps; r++){
      const double t0 = now();
      solver.makeHeff(S.gStorage(), out.data(), &S, T.Lt, T.A, T.B, T.C, T.Dd, T.S0, T.S1, T.F0, T.F1, T.Q, T.X, 0, NULL);
      const double dt = now() - t0; tot += dt; if (dt < best) best = dt;
   }
   const double t1 = now();
   solver.fillHeffDiag(diag.data(), &S, T.C, T.Dd, T.F0, T.F1, T.X, 0, NULL);
   const double t_diag = now() - t1;
   const char * ofile = args(argc, argv, "--out", NULL);
   if (ofile){ FILE * g = fopen(ofile, "wb"); fwrite(out.data(), 8, n, g); fwrite(diag.data(), 8, n, g); fclose(g); }
   double nrm = 0.0; for (long long i = 0; i < n; i++) nrm += out[i] * out[i];
   printf("B2REF synth site %d veclength %lld nkappa %d reps %d mean_s %.6f best_s %.6f diag_s %.6f setup_s %.3f threads %d norm2 %.12e\n",
          site, n, S.gNKappa(), reps, tot / reps, best, t_diag, t_setup, omp_get_max_threads(), nrm);
   return 0;
}

/* ------------------------------------------------------------------------------------------------------------------
 * "synthupdate" mode: the reference's DMRG::updateMovingRight / updateMovingLeft (DMRGoperators.cpp:243-907) on a synthetic workload at
 * full size.  A DMRG object is constructed at a tiny bond dimension (cheap PreSolve), its bookkeeper is then re-dimensioned to the
 * requested sector table, the operator tables of the OLD boundary are allocated with DMRG::allocateTensors and hash-filled (same fill
 * as `synth` / b2_opset_fill_hash), the site tensor is hash-filled, and the reference's own update routine builds every operator of the
 * NEW boundary.  Per new operator three numbers are dumped (sum, sum of squares, dot product with a hash vector): compact, and any
 * block-level difference shows. */
static void fill_dmrg_tables(DMRG & d, int t, bool mr, unsigned long long seed, double amp){
   const int L = d.L; const int side = mr ? 1 : 2;
   const int n_in = mr ? t + 1 : L - 1 - t, n_out = mr ? L - 1 - t : t + 1;
   for (int k = 0; k < n_in; k++){ const int s = mr ? t - k : t + 1 + k; fill_tensor(d.Ltensors[t][k], seed, op_key(side, K_L, s, s), amp); }
   for (int c2 = 0; c2 < n_in; c2++) for (int c3 = 0; c3 < n_in - c2; c3++){
      int i, j; if (mr){ j = t - c3; i = j - c2; } else { i = t + 1 + c3; j = i + c2; }
      fill_tensor(d.F0tensors[t][c2][c3], seed, op_key(side, K_F0, i, j), amp); fill_tensor(d.F1tensors[t][c2][c3], seed, op_key(side, K_F1, i, j), amp);
      fill_tensor(d.S0tensors[t][c2][c3], seed, op_key(side, K_S0, i, j), amp);
      if (c2 > 0) fill_tensor(d.S1tensors[t][c2][c3], seed, op_key(side, K_S1, i, j), amp);
   }
   for (int c2 = 0; c2 < n_out; c2++) for (int c3 = 0; c3 < n_out - c2; c3++){
      int i, j; if (mr){ i = t + 1 + c3; j = i + c2; } else { j = t - c3; i = j - c2; }
      fill_tensor(d.Atensors[t][c2][c3], seed, op_key(side, K_A, i, j), amp);
      if (c2 > 0) fill_tensor(d.Btensors[t][c2][c3], seed, op_key(side, K_B, i, j), amp);
      fill_tensor(d.Ctensors[t][c2][c3], seed, op_key(side, K_C, i, j), amp); fill_tensor(d.Dtensors[t][c2][c3], seed, op_key(side, K_D, i, j), amp);
   }
   for (int c2 = 0; c2 < n_out; c2++){ const int s = mr ? t + 1 + c2 : t - c2; fill_tensor(d.Qtensors[t][c2], seed, op_key(side, K_Q, s, s), amp); }
   fill_tensor(d.Xtensors[t], seed, op_key(side, K_X, -1, -1), amp);
}

static int run_synth_update(int argc, char ** argv){
   Setup s = make_setup(argc, argv);   /* --problem file */
   const int L = s.prob->gL();
   const int index = argi(argc, argv, "--site", L / 2 - 1);     /* the site whose tensor was "just optimised" */
   const bool mr = argi(argc, argv, "--moving-right", 1) != 0;
   const unsigned long long seed = (unsigned long long) argi(argc, argv, "--seed", 1);
   const double amp = argd(argc, argv, "--amp", 1.0), amp_t = argd(argc, argv, "--amp-t", 0.1);
   const double t_setup0 = now();
   /* A DMRG object WITHOUT running its constructor: DMRG::DMRG always ends with PreSolve, which for 40-60 orbitals costs a minute even at a
      tiny bond dimension.  updateMovingRight/Left, allocateTensors and deleteTensors only touch the members set below (same initial values
      as DMRG.cpp:45-100); the object is never destructed. */
   s.prob->construct_mxelem();
   DMRG & d = *static_cast<DMRG *>(calloc(1, sizeof(DMRG)));
   d.Prob = s.prob; d.L = L; d.nStates = 1; d.Exc_activated = false; d.makecheckpoints = false;
   d.denBK = new SyBookkeeper(s.prob, 2);
   d.Ltensors = new TensorL ** [L - 1]; d.F0tensors = new TensorF0 *** [L - 1]; d.F1tensors = new TensorF1 *** [L - 1];
   d.S0tensors = new TensorS0 *** [L - 1]; d.S1tensors = new TensorS1 *** [L - 1];
   d.Atensors = new TensorOperator *** [L - 1]; d.Btensors = new TensorOperator *** [L - 1]; d.Ctensors = new TensorOperator *** [L - 1]; d.Dtensors = new TensorOperator *** [L - 1];
   d.Qtensors = new TensorQ ** [L - 1]; d.Xtensors = new TensorX * [L - 1];
   d.isAllocated = new int[L - 1];
   for (int cnt = 0; cnt < L - 1; cnt++) d.isAllocated[cnt] = 0;
   const char * dfile = args(argc, argv, "--dims", NULL);
   if (!dfile){ fprintf(stderr, "synthupdate: --dims file needed\n"); return 1; }
   { FILE * g = fopen(dfile, "rb"); if (!g){ perror(dfile); return 2; }
     int row[5]; while (fread(row, 4, 5, g) == 5) d.denBK->SetDim(row[0], row[1], row[2], row[3], row[4]);
     fclose(g); }
   d.MPS = new TensorT * [L];
   for (int site = 0; site < L; site++) d.MPS[site] = new TensorT(site, d.denBK);

   const int t_new = mr ? index : index - 1;        /* table slot that receives the new operators */
   const int t_old = mr ? index - 1 : index;        /* slot of the operators one site further out  */
   const bool have_old = mr ? (index > 0) : (index < L - 1);
   if (t_new < 0 || t_new > L - 2){ fprintf(stderr, "synthupdate: no operators live there\n"); return 1; }
   if (have_old){ d.allocateTensors(t_old, mr); d.isAllocated[t_old] = mr ? 1 : 2; fill_dmrg_tables(d, t_old, mr, seed, amp); }
   d.allocateTensors(t_new, mr); d.isAllocated[t_new] = mr ? 1 : 2;
   { TensorT * T = d.MPS[index]; const long long n = T->gKappa2index(T->gNKappa()); double * p = T->gStorage();
     for (long long e = 0; e < n; e++) p[e] = amp_t * hash_value(seed, op_key(4, 0, -1, -1), e); }
   const double t_setup = now() - t_setup0;
   const double t0 = now();
   if (mr) d.updateMovingRight(t_new); else d.updateMovingLeft(t_new);
   const double dt = now() - t0;
   std::vector<int> meta; std::vector<double> sums;
   { std::vector<int> m2; std::vector<double> data;   /* walk the tables of slot t_new in dump_ops order without copying the data */
     struct V { std::vector<int> & meta; std::vector<double> & sums; unsigned long long seed;
                void add(int kind, int i, int j, Tensor * t){ if (!t) return; const long long n = t->gKappa2index(t->gNKappa()); const double * p = t->gStorage();
                   double a = 0.0, b = 0.0, c = 0.0;
                   for (long long e = 0; e < n; e++){ a += p[e]; b += p[e] * p[e]; c += p[e] * hash_value(seed + 17, op_key(5, kind, i, j), e); }
                   meta.push_back(kind); meta.push_back(i); meta.push_back(j); meta.push_back((int) n); sums.push_back(a); sums.push_back(b); sums.push_back(c); } } v = { meta, sums, seed };
     const int t = t_new;
     const int n_in = mr ? t + 1 : L - 1 - t, n_out = mr ? L - 1 - t : t + 1;
     for (int k = 0; k < n_in; k++){ const int st = mr ? t - k : t + 1 + k; v.add(K_L, st, st, d.Ltensors[t][k]); }
     for (int c2 = 0; c2 < n_in; c2++) for (int c3 = 0; c3 < n_in - c2; c3++){
        int i, j; if (mr){ j = t - c3; i = j - c2; } else { i = t + 1 + c3; j = i + c2; }
        v.add(K_S0, i, j, d.S0tensors[t][c2][c3]); if (c2 > 0) v.add(K_S1, i, j, d.S1tensors[t][c2][c3]);
        v.add(K_F0, i, j, d.F0tensors[t][c2][c3]); v.add(K_F1, i, j, d.F1tensors[t][c2][c3]);
     }
     for (int c2 = 0; c2 < n_out; c2++) for (int c3 = 0; c3 < n_out - c2; c3++){
        int i, j; if (mr){ i = t + 1 + c3; j = i + c2; } else { j = t - c3; i = j - c2; }
        v.add(K_A, i, j, d.Atensors[t][c2][c3]); if (c2 > 0) v.add(K_B, i, j, d.Btensors[t][c2][c3]);
        v.add(K_C, i, j, d.Ctensors[t][c2][c3]); v.add(K_D, i, j, d.Dtensors[t][c2][c3]);
     }
     for (int c2 = 0; c2 < n_out; c2++){ const int st = mr ? t + 1 + c2 : t - c2; v.add(K_Q, st, st, d.Qtensors[t][c2]); }
     v.add(K_X, -1, -1, d.Xtensors[t]);
   }
   Writer w(args(argc, argv, "--out", "synthupdate.b2fx"));
   w.ints("upd/meta", meta); w.dbls("upd/sums", sums);
   printf("B2REF synthupdate site %d moving_right %d operators %d update_s %.6f setup_s %.3f threads %d\n", index, mr ? 1 : 0, (int)(meta.size() / 4), dt, t_setup, omp_get_max_threads());
   return 0;
}

int main(int argc, char ** argv){
   if (argc < 2){ fprintf(stderr, "usage: ref_driver dump|energies|time|wigner ...\n"); return 1; }
   const std::string mode = argv[1];
   std::cout.precision(15);

   if (mode == "wigner"){   /* table of 6j / 9j values for the parity test of b2::wigner6j/9j */
      Writer w(args(argc, argv, "--out", "wigner.b2fx"));
      std::vector<int> a6; std::vector<double> v6; std::vector<int> a9; std::vector<double> v9;
      unsigned int st = 777u;
      for (int n = 0; n < 4000; n++){
         int j[9]; for (int k = 0; k < 9; k++){ st = st * 1664525u + 1013904223u; j[k] = (st >> 10) % (n < 2000 ? 7 : 40); }
         a6.insert(a6.end(), j, j + 6); v6.push_back(Wigner::wigner6j(j[0], j[1], j[2], j[3], j[4], j[5]));
         for (int k = 0; k < 9; k++) j[k] = j[k] % 9;
         a9.insert(a9.end(), j, j + 9); v9.push_back(Wigner::wigner9j(j[0], j[1], j[2], j[3], j[4], j[5], j[6], j[7], j[8]));
      }
      w.ints("w6j/args", a6); w.dbls("w6j/vals", v6); w.ints("w9j/args", a9); w.dbls("w9j/vals", v9);
      return 0;
   }

   if (mode == "synth") return run_synth(argc, argv);
   if (mode == "synthupdate") return run_synth_update(argc, argv);
   if (mode == "problem"){   /* only the problem (orbital irreps in DMRG order + folded integral table): input of full calculations */
      Setup s = make_setup(argc, argv);
      s.prob->construct_mxelem();
      Writer w(args(argc, argv, "--out", "problem.b2fx"));
      dump_problem(w, s.prob, s.ham, s.group);
      printf("B2REF problem dumped\n");
      return 0;
   }

   Setup s = make_setup(argc, argv);
   const int D = argi(argc, argv, "--D", 20);
   const int seed = argi(argc, argv, "--seed", 1234);
   const double rtol = argd(argc, argv, "--rtol", 1e-8);
   const double noise = argd(argc, argv, "--noise", 0.0);
   const int L = s.prob->gL();

   if (mode == "energies"){   /* schedule "D:econv:maxsweeps:noise:rtol,..." */
      std::string sched = args(argc, argv, "--schedule", "");
      std::vector<std::vector<double> > ins;
      { std::stringstream ss(sched); std::string item;
        while (std::getline(ss, item, ',')){ std::vector<double> v; std::stringstream s2(item); std::string x; while (std::getline(s2, x, ':')) v.push_back(atof(x.c_str())); ins.push_back(v); } }
      ConvergenceScheme scheme(ins.size());
      for (size_t i = 0; i < ins.size(); i++) scheme.set_instruction(i, (int) ins[i][0], ins[i][1], (int) ins[i][2], ins[i][3], ins[i][4]);
      srand(seed);
      const double t0 = now();
      /* --chkpt-dir DIR: run inside DIR with MPS checkpoints on (DMRG.cpp:57-65,105-116,334): CheMPS2_MPS0.h5 there is loaded by the
         constructor when it exists (resume) and rewritten after every sweep — through env_shims/hdf5.h the container chemps2_b200's
         b2_dmrg_save_mps / _load_mps use, so the reference and the GPU library resume each other's states */
      const char * chk = args(argc, argv, "--chkpt-dir", NULL);
      if (chk && chdir(chk) != 0){ perror(chk); return 2; }
      DMRG d(s.prob, &scheme, chk != NULL, chk ? chk : "/tmp");
      apply_pairing_model(s, d);
      apply_momentum_hubbard(s, d);
      const double e = d.Solve();
      printf("B2REF final_energy %.15f wall %.3f threads %d\n", e, now() - t0, omp_get_max_threads());
      return 0;
   }

   ConvergenceScheme scheme(1);
   scheme.set_instruction(0, D, 1e-10, 2, noise, rtol);
   srand(seed);
   DMRG d(s.prob, &scheme, false, "/tmp");   /* random MPS + PreSolve (all right-moving operators) */
   apply_pairing_model(s, d);
   apply_momentum_hubbard(s, d);

   if (mode == "dump"){
      Writer w(args(argc, argv, "--out", "case.b2fx"));
      const int siteA = argi(argc, argv, "--siteA", L / 2);       /* sigma case met during the left sweep  */
      const int siteB = argi(argc, argv, "--siteB", L / 2 - 1);   /* sigma case met during the right sweep */
      const int presweeps = argi(argc, argv, "--presweeps", 0);   /* full left+right sweeps before the dumps */
      dump_problem(w, s.prob, s.ham, s.group);
      std::vector<double> energies;
      bool change = false;
      for (int sw = 0; sw < presweeps; sw++){
         energies.push_back(d.sweepleft(change, 0, true)); change = true;
         energies.push_back(d.sweepright(change, 0, true));
      }
      /* left sweep by hand (DMRG.cpp:357-386) */
      for (int index = L - 2; index > 0; index--){
         if (index == siteA){
            dump_bk(w, "A/bk", d.denBK); dump_mps(w, "A/mps", d); dump_sigma(w, "A", d, index);
         }
         const double e = d.solve_site(index, rtol, 0.0, D, true, false, change);
         energies.push_back(e);
         if (index == siteA){   /* operator update moving left: inputs = A/right ops + new MPS[index+1]; outputs = table `index` */
            dump_bk(w, "UL/bk", d.denBK); dump_mps(w, "UL/mps", d);
         }
         if (index == siteA){
            /* TwoDM::FillSite at this site (TwoDM.cpp:445-628): T = MPS[siteA], left operators of boundary siteA (A/left, table siteA-1),
               right operators of boundary siteA+1 (UL/new, table siteA).  updateMovingLeftSafe frees table siteA-1 at its end, so the
               right table is built first by the same steps updateMovingLeftSafe starts with (DMRGoperators.cpp:124-132); the Safe call
               below then only repeats updateMovingLeft and does the life-cycle bookkeeping. */
            if (d.isAllocated[index] == 1){ d.deleteTensors(index, true); d.isAllocated[index] = 0; }
            if (d.isAllocated[index] == 0){ d.allocateTensors(index, false); d.isAllocated[index] = 2; }
            d.updateMovingLeft(index);
            TwoDM tdm(d.denBK, d.Prob);
            tdm.FillSite(d.MPS[index], d.Ltensors, d.F0tensors, d.F1tensors, d.S0tensors, d.S1tensors);
            const long long n4 = (long long) L * L * L * L;
            w.dbls("UL/twodm_A", tdm.two_rdm_A, n4);
            w.dbls("UL/twodm_B", tdm.two_rdm_B, n4);
         }
         d.updateMovingLeftSafe(index);
         if (index == siteA) dump_ops(w, "UL/new", d, index, false);
      }
      change = true;
      for (int index = 0; index < L - 2; index++){
         if (index == siteB){
            dump_bk(w, "B/bk", d.denBK); dump_mps(w, "B/mps", d); dump_sigma(w, "B", d, index);
         }
         const double e = d.solve_site(index, rtol, 0.0, D, true, true, change);
         energies.push_back(e);
         if (index == siteB){ dump_bk(w, "UR/bk", d.denBK); dump_mps(w, "UR/mps", d); }
         d.updateMovingRightSafe(index);
         if (index == siteB) dump_ops(w, "UR/new", d, index, true);
      }
      w.dbls("energies", energies);
      dump_correlation_tensors(w, "corr", d);
      /* the full 2-RDM of the final MPS (= corr/mps): DMRG::calc2DMandCorrelations -> TwoDM arrays A and B in DMRG orbital order */
      d.calc2DMandCorrelations();
      {
         const long long n4 = (long long) L * L * L * L;
         w.dbls("twodm/A", d.the2DM->two_rdm_A, n4);
         w.dbls("twodm/B", d.the2DM->two_rdm_B, n4);
         std::vector<double> te; te.push_back(d.the2DM->trace()); te.push_back(d.the2DM->energy());
         w.dbls("twodm/trace_energy", te);
         const long long n2 = (long long) L * L;   /* Correlations tables of the same state (Correlations.cpp), DMRG orbital order */
         w.dbls("corrfun/Cspin", d.theCorr->Cspin, n2); w.dbls("corrfun/Cdens", d.theCorr->Cdens, n2);
         w.dbls("corrfun/Cspinflip", d.theCorr->Cspinflip, n2); w.dbls("corrfun/Cdirad", d.theCorr->Cdirad, n2);
         w.dbls("corrfun/MutInfo", d.theCorr->MutInfo, n2);
      }
      printf("B2REF dumped; last energy %.12f\n", energies.back());
      return 0;
   }

   if (mode == "trace"){
      /* Step-by-step record of sweeps started from the seeded random MPS with noise: solve_site (DMRG.cpp:419-452) spelled out through the
         public methods so that the input and every output of Sobject::Split (Sobject.cpp:260-622) can be dumped: the two-site object after
         Davidson + Sobject::addNoise (rand() stream continued from TensorT::random), the discarded weight, the re-dimensioned boundary, the
         new site tensors and their Join (gauge-invariant).  Half sweeps: left (fixed dimensions), right, left. */
      Writer w(args(argc, argv, "--out", "trace.b2fx"));
      dump_problem(w, s.prob, s.ham, s.group);
      dump_bk(w, "trace/bk0", d.denBK); dump_mps(w, "trace/mps0", d);
      const int halfsweeps = argi(argc, argv, "--halfsweeps", 3);
      std::vector<int> hdr; hdr.push_back(D); hdr.push_back(seed); std::vector<double> pars; pars.push_back(rtol); pars.push_back(noise);
      int step = 0;
      for (int hs = 0; hs < halfsweeps; hs++){
         const bool mr = (hs % 2 == 1); const bool change = (hs > 0);
         for (int k = 0; k < L - 2 + (mr ? 0 : -1) + (mr ? 0 : 1); k++){
            const int index = mr ? k : L - 2 - k;
            if (!mr && index == 0) break;
            std::ostringstream nm; nm << "trace/s" << step;
            Sobject * denS = new Sobject(index, d.denBK);
            denS->Join(d.MPS[index], d.MPS[index + 1]);
            Heff Solver(d.denBK, d.Prob, rtol);
            double E = Solver.SolveDAVIDSON(denS, d.Ltensors, d.Atensors, d.Btensors, d.Ctensors, d.Dtensors, d.S0tensors, d.S1tensors, d.F0tensors,
                                            d.F1tensors, d.Qtensors, d.Xtensors, 0, NULL) + s.prob->gEconst();
            if (noise > 0.0) denS->addNoise(noise);
            dump_bk(w, nm.str() + "/bk", d.denBK);
            w.dbls(nm.str() + "/S", denS->gStorage(), denS->gKappa2index(denS->gNKappa()));
            const double dw = denS->Split(d.MPS[index], d.MPS[index + 1], D, mr, change);
            delete denS;
            std::vector<int> sh; sh.push_back(index); sh.push_back(mr ? 1 : 0); sh.push_back(change ? 1 : 0); sh.push_back(D);
            w.ints(nm.str() + "/hdr", sh);
            std::vector<double> res; res.push_back(E); res.push_back(dw); w.dbls(nm.str() + "/res", res);
            dump_bk(w, nm.str() + "/bk_after", d.denBK);
            w.dbls(nm.str() + "/tl", d.MPS[index]->gStorage(), d.MPS[index]->gKappa2index(d.MPS[index]->gNKappa()));
            w.dbls(nm.str() + "/tr", d.MPS[index + 1]->gStorage(), d.MPS[index + 1]->gKappa2index(d.MPS[index + 1]->gNKappa()));
            { Sobject J(index, d.denBK); J.Join(d.MPS[index], d.MPS[index + 1]); w.dbls(nm.str() + "/joined_after", J.gStorage(), J.gKappa2index(J.gNKappa())); }
            if (mr) d.updateMovingRightSafe(index); else d.updateMovingLeftSafe(index);
            step++;
         }
      }
      hdr.push_back(step); w.ints("trace/hdr", hdr); w.dbls("trace/pars", pars);
      printf("B2REF trace: %d steps\n", step);
      return 0;
   }

   if (mode == "time"){   /* time Heff::makeHeff at one site; operators are whatever the un-optimised random MPS gives */
      const int site = argi(argc, argv, "--site", L / 2);
      const int reps = argi(argc, argv, "--reps", 3);
      for (int index = L - 2; index > site; index--) d.updateMovingLeftSafe(index);   /* no solve: only build right operators */
      Sobject S(site, d.denBK);
      S.Join(d.MPS[site], d.MPS[site + 1]);
      S.prog2symm();
      const int n = S.gKappa2index(S.gNKappa());
      std::vector<double> out(n);
      Heff solver(d.denBK, d.Prob, 1e-5);
      double best = 1e99, tot = 0.0;
      for (int r = 0; r < reps + 1; r++){
         const double t0 = now();
         solver.makeHeff(S.gStorage(), out.data(), &S, d.Ltensors, d.Atensors, d.Btensors, d.Ctensors, d.Dtensors, d.S0tensors, d.S1tensors,
                         d.F0tensors, d.F1tensors, d.Qtensors, d.Xtensors, 0, NULL);
         const double dt = now() - t0;
         if (r > 0){ tot += dt; if (dt < best) best = dt; }
      }
      double nrm = 0.0; for (int i = 0; i < n; i++) nrm += out[i] * out[i];
      printf("B2REF time site %d veclength %d nkappa %d reps %d mean_s %.6f best_s %.6f threads %d norm2 %.12e\n",
             site, n, S.gNKappa(), reps, tot / reps, best, omp_get_max_threads(), nrm);
      return 0;
   }
   fprintf(stderr, "unknown mode\n");
   return 1;
}
