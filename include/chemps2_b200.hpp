/* chemps2_b200.hpp — header-only C++ mirror of the reference's caller-facing classes for the two-site DMRG path, on top of the C ABI
 * of chemps2_b200.h.  A program written against CheMPS2's public headers for this path (tests/test1..5,12.cpp.in, executable.cpp,
 * PyCheMPS2's DMRGsolver.pxd) compiles against this header with the same class names, method names, argument meaning and error
 * behaviour, and runs its sweeps on the GPU:
 *
 *    reference class (file:line)                      mirror below
 *    Irreps            (Irreps.h:74-127)               Irreps            group names, directProd, number of irreps, psi4 <-> molpro labels
 *    Hamiltonian       (Hamiltonian.h:61-169)          Hamiltonian       orbitals, irreps, Econst, Tmat, Vmat (8-fold symmetric), FCIDUMP reader
 *    Problem           (Problem.h:44-132)              Problem           target sector, orbital reordering f1/f2, folded table gMxElement
 *    ConvergenceScheme (ConvergenceScheme.h:47-98)     ConvergenceScheme D / Econv / max sweeps / noise prefactor / Davidson rtol per instruction
 *    DMRG              (DMRG.h:93-162)                 DMRG              PreSolve, Solve, calc2DMandCorrelations, get2DM, getCorrelations,
 *                                                                        activateExcitations, newExcitation, deleteStoredMPS/Operators
 *    TwoDM             (TwoDM.h:57-136)                TwoDM             getTwoDMA/B_DMRG/HAM, get1RDM_*, spin_density_*, trace, energy
 *    Correlations      (Correlations.h:106-193)        Correlations      Cspin/Cdens/Cspinflip/Cdirad/MutualInformation, entropies
 *
 * Error behaviour: the reference asserts on impossible input and crashes; the mirror prints the library's error string and aborts.
 * There is no CPU fallback: constructing a DMRG object without a CUDA device aborts with the library's message.
 * Not mirrored (outside the path, DESIGN.md "out of scope"): CASSCF, EdmistonRuedenberg, FCI, ThreeDM, Molden, HDF5 files.
 * The namespace defaults to CheMPS2 so existing callers compile unchanged; define CHEMPS2_B200_NAMESPACE to rename it when a
 * program also links the reference library. */
#ifndef CHEMPS2_B200_HPP
#define CHEMPS2_B200_HPP

#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <sys/time.h>
#include <vector>

#include "chemps2_b200.h"

#ifndef CHEMPS2_B200_NAMESPACE
#define CHEMPS2_B200_NAMESPACE CheMPS2
#endif

namespace CHEMPS2_B200_NAMESPACE {

const double DAVIDSON_DMRG_RTOL = 1e-5;               /* Options.h:73 */
const double CORRELATIONS_discardEig = 1e-100;        /* Options.h:97 */
const bool DMRG_storeMpsOnDisk = false;               /* Options.h:41 */
const bool DMRG_storeRenormOptrOnDisk = false;        /* Options.h:40: operators stay in HBM / pinned host memory */
const std::string defaultTMPpath = "/tmp";            /* Options.h:30 */
const std::string DMRG_MPS_storage_prefix = "CheMPS2_MPS";

namespace detail {
inline void check(int rc, const char* what) {
   if (rc == 0) return;
   std::fprintf(stderr, "chemps2_b200: %s failed (code %d): %s\n", what, rc, b2_last_error());
   std::abort();
}
inline double now() { struct timeval t; gettimeofday(&t, NULL); return t.tv_sec + 1e-6 * t.tv_usec; }
}

/* seed of the random initial MPS; Initialize::Init (Initialize.cpp:29) seeds rand() from the clock, here the stream is the library's own
 * and the seed is explicit (fixed default: runs are reproducible) */
class Initialize {
public:
   static unsigned long long& seed() { static unsigned long long s = 12345ULL; return s; }
   static void Init() {}
   static void SetSeed(const unsigned long long s) { seed() = s; }
};

class Irreps {
public:
   Irreps() : isActivated(false), groupNumber(0) {}
   Irreps(const int nGroup) : isActivated(false), groupNumber(0) { setGroup(nGroup); }
   bool setGroup(const int nGroup) {
      if (nGroup >= 0 && nGroup <= 7) { isActivated = true; groupNumber = nGroup; }
      return isActivated;
   }
   bool getIsActivated() const { return isActivated; }
   int getGroupNumber() const { return isActivated ? groupNumber : -1; }
   std::string getGroupName() const { return isActivated ? getGroupName(groupNumber) : "error1"; }
   static std::string getGroupName(const int nGroup) {
      static const char* names[8] = {"c1", "ci", "c2", "cs", "d2", "c2v", "c2h", "d2h"};
      return (nGroup >= 0 && nGroup <= 7) ? names[nGroup] : "error2";
   }
   int getNumberOfIrreps() const { return isActivated ? getNumberOfIrreps(groupNumber) : -1; }
   static int getNumberOfIrreps(const int nGroup) {
      static const int num[8] = {1, 2, 2, 2, 4, 4, 4, 8};
      return (nGroup >= 0 && nGroup <= 7) ? num[nGroup] : -1;
   }
   /* psi4 numbering: the product of two irreps is the XOR of their numbers (Irreps.h:111) */
   static int directProd(const int Irrep1, const int Irrep2) { return Irrep1 ^ Irrep2; }
   /* molpro irrep number (1-based, as in FCIDUMP ORBSYM) of every psi4 irrep of the group (Irreps.cpp:184-216) */
   void symm_psi2molpro(int* psi2molpro) const { if (isActivated) symm_psi2molpro(psi2molpro, getGroupName()); }
   static void symm_psi2molpro(int* psi2molpro, const std::string SymmLabel) {
      static const int two[2] = {1, 2}, d2[4] = {1, 4, 3, 2}, c2v[4] = {1, 4, 2, 3}, d2h[8] = {1, 4, 6, 7, 8, 5, 3, 2};
      const int* src = NULL; int n = 0;
      if (SymmLabel == "c1") { src = two; n = 1; }
      else if (SymmLabel == "ci" || SymmLabel == "c2" || SymmLabel == "cs") { src = two; n = 2; }
      else if (SymmLabel == "d2") { src = d2; n = 4; }
      else if (SymmLabel == "c2v" || SymmLabel == "c2h") { src = c2v; n = 4; }
      else if (SymmLabel == "d2h") { src = d2h; n = 8; }
      for (int i = 0; i < n; i++) psi2molpro[i] = src[i];
   }
private:
   bool isActivated;
   int groupNumber;
};

/* Hamiltonian.h:61-169.  Storage is dense (L^2 and L^4 doubles) instead of the irrep-blocked TwoIndex / FourIndex: the hot path
 * reads it once, when Problem folds it into the gMxElement table. */
class Hamiltonian {
public:
   Hamiltonian(const int Norbitals, const int nGroup, const int* OrbIrreps) : L(Norbitals), SymmInfo(nGroup), Econst(0.0) {
      assert(SymmInfo.getIsActivated());
      orb2irrep.assign(OrbIrreps, OrbIrreps + L);
      for (int i = 0; i < L; i++) assert(orb2irrep[i] >= 0 && orb2irrep[i] < SymmInfo.getNumberOfIrreps());
      Tmat.assign((size_t)L * L, 0.0);
      Vmat.assign((size_t)L * L * L * L, 0.0);
   }
   /* FCIDUMP with molpro ORBSYM labels, converted to psi4 numbering for group psi4groupnumber (Hamiltonian.cpp:296-420) */
   Hamiltonian(const std::string filename, const int psi4groupnumber) : L(0), SymmInfo(psi4groupnumber), Econst(0.0) {
      assert(SymmInfo.getIsActivated());
      CreateAndFillFromFCIDUMP(filename);
   }
   virtual ~Hamiltonian() {}
   int getL() const { return L; }
   int getNGroup() const { return SymmInfo.getGroupNumber(); }
   int getOrbitalIrrep(const int nOrb) const { return orb2irrep[nOrb]; }
   void setEconst(const double val) { Econst = val; }
   double getEconst() const { return Econst; }
   void setTmat(const int index1, const int index2, const double val) {
      assert(orb2irrep[index1] == orb2irrep[index2]);
      Tmat[index1 + (size_t)L * index2] = val;
      Tmat[index2 + (size_t)L * index1] = val;
   }
   double getTmat(const int index1, const int index2) const {
      return orb2irrep[index1] == orb2irrep[index2] ? Tmat[index1 + (size_t)L * index2] : 0.0;
   }
   /* physicist notation <12|34>; all eight permutation-equivalent elements are written together, like FourIndex::set */
   void setVmat(const int i1, const int i2, const int i3, const int i4, const double val) {
      assert(Irreps::directProd(orb2irrep[i1], orb2irrep[i2]) == Irreps::directProd(orb2irrep[i3], orb2irrep[i4]));
      eightfold(i1, i2, i3, i4, val, false);
   }
   void addToVmat(const int i1, const int i2, const int i3, const int i4, const double val) {
      assert(Irreps::directProd(orb2irrep[i1], orb2irrep[i2]) == Irreps::directProd(orb2irrep[i3], orb2irrep[i4]));
      eightfold(i1, i2, i3, i4, vat(i1, i2, i3, i4) + val, false);
   }
   double getVmat(const int i1, const int i2, const int i3, const int i4) const {
      if (Irreps::directProd(orb2irrep[i1], orb2irrep[i2]) != Irreps::directProd(orb2irrep[i3], orb2irrep[i4])) return 0.0;
      return vat(i1, i2, i3, i4);
   }
   void writeFCIDUMP(const std::string fcidumpfile, const int Nelec, const int TwoS, const int TargetIrrep) const {
      std::vector<int> p2m(SymmInfo.getNumberOfIrreps());
      SymmInfo.symm_psi2molpro(p2m.data());
      FILE* f = std::fopen(fcidumpfile.c_str(), "w");
      assert(f != NULL);
      std::fprintf(f, " &FCI NORB= %d,NELEC= %d,MS2= %d,\n  ORBSYM=", L, Nelec, TwoS);
      for (int i = 0; i < L; i++) std::fprintf(f, "%d,", p2m[orb2irrep[i]]);
      std::fprintf(f, "\n  ISYM=%d,\n /\n", p2m[TargetIrrep]);
      for (int p = 0; p < L; p++) for (int q = 0; q <= p; q++) {            /* chemist (pq|rs), unique elements only */
         const int ipq = Irreps::directProd(orb2irrep[p], orb2irrep[q]);
         for (int r = 0; r <= p; r++) for (int s = 0; s <= r; s++) {
            if (r == p && s > q) continue;
            if (Irreps::directProd(orb2irrep[r], orb2irrep[s]) != ipq) continue;
            std::fprintf(f, " % 23.16E %3d %3d %3d %3d\n", getVmat(p, r, q, s), p + 1, q + 1, r + 1, s + 1);
         }
      }
      for (int p = 0; p < L; p++) for (int q = 0; q <= p; q++)
         if (orb2irrep[p] == orb2irrep[q]) std::fprintf(f, " % 23.16E %3d %3d %3d %3d\n", getTmat(p, q), p + 1, q + 1, 0, 0);
      std::fprintf(f, " % 23.16E %3d %3d %3d %3d\n", Econst, 0, 0, 0, 0);
      std::fclose(f);
   }
private:
   int L;
   Irreps SymmInfo;
   double Econst;
   std::vector<int> orb2irrep;
   std::vector<double> Tmat, Vmat;
   size_t vidx(int a, int b, int c, int d) const { return a + (size_t)L * (b + (size_t)L * (c + (size_t)L * d)); }
   double vat(int a, int b, int c, int d) const { return Vmat[vidx(a, b, c, d)]; }
   void eightfold(int a, int b, int c, int d, double v, bool) {
      Vmat[vidx(a, b, c, d)] = v; Vmat[vidx(c, b, a, d)] = v; Vmat[vidx(a, d, c, b)] = v; Vmat[vidx(c, d, a, b)] = v;
      Vmat[vidx(b, a, d, c)] = v; Vmat[vidx(d, a, b, c)] = v; Vmat[vidx(b, c, d, a)] = v; Vmat[vidx(d, c, b, a)] = v;
   }
   void CreateAndFillFromFCIDUMP(const std::string fcidumpfile) {
      struct stat info;
      const bool on_disk = fcidumpfile.length() > 0 && stat(fcidumpfile.c_str(), &info) == 0;
      if (!on_disk) std::cout << "CheMPS2::Hamiltonian : Unable to find FCIDUMP file " << fcidumpfile << "!" << std::endl;
      assert(on_disk);
      std::ifstream in(fcidumpfile.c_str());
      std::string header, line;
      while (std::getline(in, line)) {                           /* namelist up to the closing "/" or "&END" */
         std::string t = line;
         t.erase(0, t.find_first_not_of(" \t"));
         if (t.compare(0, 1, "/") == 0 || t.find("&END") != std::string::npos || t.find("&end") != std::string::npos) break;
         header += line + " ";
      }
      size_t pos = header.find("NORB");
      assert(pos != std::string::npos);
      L = std::atoi(header.c_str() + header.find("=", pos) + 1);
      assert(L > 0);
      const int nIrreps = SymmInfo.getNumberOfIrreps();
      std::vector<int> p2m(nIrreps);
      SymmInfo.symm_psi2molpro(p2m.data());
      orb2irrep.assign(L, -1);
      pos = header.find("ORBSYM");
      assert(pos != std::string::npos);
      pos = header.find("=", pos) + 1;
      for (int orb = 0; orb < L; orb++) {
         while (pos < header.size() && (header[pos] == ' ' || header[pos] == ',')) pos++;
         const int molpro = std::atoi(header.c_str() + pos);
         while (pos < header.size() && header[pos] != ',' && header[pos] != ' ') pos++;
         for (int ir = 0; ir < nIrreps; ir++) if (p2m[ir] == molpro) orb2irrep[orb] = ir;
         assert(orb2irrep[orb] != -1);
      }
      Tmat.assign((size_t)L * L, 0.0);
      Vmat.assign((size_t)L * L * L * L, 0.0);
      double value; int i1, i2, i3, i4;
      while (in >> value >> i1 >> i2 >> i3 >> i4) {
         if (i4 != 0) setVmat(i1 - 1, i3 - 1, i2 - 1, i4 - 1, value);      /* chemist (12|34) -> physicist <13|24> */
         else if (i2 != 0) setTmat(i1 - 1, i2 - 1, value);
         else { Econst = value; break; }
      }
   }
};

/* Problem.h:44-132 */
class Problem {
public:
   Problem(const Hamiltonian* Hamin, const int TwoSin, const int Nin, const int Irrepin)
      : Ham(Hamin), L(Hamin->getL()), TwoS(TwoSin), N(Nin), Irrep(Irrepin), bReorder(false) { checkConsistency(); }
   virtual ~Problem() {}
   int gL() const { return L; }
   int gSy() const { return Ham->getNGroup(); }
   int gIrrep(const int nOrb) const { return Ham->getOrbitalIrrep(bReorder ? f2[nOrb] : nOrb); }
   int gTwoS() const { return TwoS; }
   int gN() const { return N; }
   int gIrrep() const { return Irrep; }
   double gEconst() const { return Ham->getEconst(); }
   double gMxElement(const int a, const int b, const int c, const int d) const { return mx_elem[a + (size_t)L * (b + (size_t)L * (c + (size_t)L * d))]; }
   void setMxElement(const int a, const int b, const int c, const int d, const double v) { mx_elem[a + (size_t)L * (b + (size_t)L * (c + (size_t)L * d))] = v; }
   /* V + (T spread over the N-1 partner electrons), in DMRG orbital order (Problem.cpp:363-384) */
   void construct_mxelem() {
      mx_elem.resize((size_t)L * L * L * L);
      const double prefact = 1.0 / (N - 1);
      for (int o1 = 0; o1 < L; o1++) { const int m1 = bReorder ? f2[o1] : o1;
         for (int o2 = 0; o2 < L; o2++) { const int m2 = bReorder ? f2[o2] : o2;
            for (int o3 = 0; o3 < L; o3++) { const int m3 = bReorder ? f2[o3] : o3;
               for (int o4 = 0; o4 < L; o4++) { const int m4 = bReorder ? f2[o4] : o4;
                  setMxElement(o1, o2, o3, o4, Ham->getVmat(m1, m2, m3, m4) + prefact * (o1 == o3 ? Ham->getTmat(m2, m4) : 0.0)
                                                   + prefact * (o2 == o4 ? Ham->getTmat(m1, m3) : 0.0));
               } } } }
   }
   const double* mx_table() const { return mx_elem.data(); }
   bool checkConsistency() const {
      Irreps SymmInfo(gSy());
      if (gIrrep() < 0 || gIrrep() >= SymmInfo.getNumberOfIrreps()) { std::cout << "Problem::Problem() : Irrep out of bound : Irrep = " << gIrrep() << std::endl; return false; }
      if (gTwoS() < 0) { std::cout << "Problem::checkConsistency() : TwoS = " << gTwoS() << std::endl; return false; }
      if (gN() < 0) { std::cout << "Problem::checkConsistency() : N = " << gN() << std::endl; return false; }
      if (gL() < 0) { std::cout << "Problem::checkConsistency() : L = " << gL() << std::endl; return false; }
      if (gN() > 2 * gL()) { std::cout << "Problem::checkConsistency() : N > 2*L ; N = " << gN() << " and L = " << gL() << std::endl; return false; }
      if ((gN() % 2) != (gTwoS() % 2)) { std::cout << "Problem::checkConsistency() : N % 2 != TwoS % 2 ; N = " << gN() << " and TwoS = " << gTwoS() << std::endl; return false; }
      if (gTwoS() > gL() - std::abs(gN() - gL())) { std::cout << "Problem::checkConsistency() : TwoS > L - |N-L| ; N = " << gN() << " and TwoS = " << gTwoS() << " and L = " << gL() << std::endl; return false; }
      return true;
   }
   bool gReorder() const { return bReorder; }
   int gf1(const int HamOrb) const { return bReorder ? f1[HamOrb] : -1; }
   int gf2(const int DMRGOrb) const { return bReorder ? f2[DMRGOrb] : -1; }
   /* D2h: sigma, sigma*, pi_x, pi_x*, pi_y, pi_y*, then B1g, Au (Problem.cpp:57-94) */
   void SetupReorderD2h() {
      bReorder = false;
      if (gSy() != 7) return;
      static const int order[8] = {0, 5, 7, 2, 6, 3, 1, 4};
      by_irrep_order(order, 8, false);
   }
   /* C2v: A1 (reversed), B1, B2, A2 (Problem.cpp:96-147) */
   void SetupReorderC2v() {
      bReorder = false;
      if (gSy() != 5) return;
      static const int order[4] = {0, 2, 3, 1};
      by_irrep_order(order, 4, true);
   }
   void setup_reorder_custom(int* dmrg2ham) {
      bReorder = true;
      f1.assign(L, -2); f2.assign(L, 0);
      for (int d = 0; d < L; d++) { assert(dmrg2ham[d] >= 0 && dmrg2ham[d] < L); f2[d] = dmrg2ham[d]; f1[dmrg2ham[d]] = d; }
      for (int h = 0; h < L; h++) assert(f1[h] >= 0);
   }
private:
   const Hamiltonian* Ham;
   int L, TwoS, N, Irrep;
   bool bReorder;
   std::vector<int> f1, f2;         /* f1[HamOrb] = DMRGOrb, f2[DMRGOrb] = HamOrb */
   std::vector<double> mx_elem;
   void by_irrep_order(const int* order, int n, bool reverse_first) {
      bReorder = true;
      f1.assign(L, 0); f2.assign(L, 0);
      int DMRGOrb = 0;
      for (int k = 0; k < n; k++) {
         if (k == 0 && reverse_first) { for (int h = L - 1; h >= 0; h--) if (Ham->getOrbitalIrrep(h) == order[k]) { f1[h] = DMRGOrb; f2[DMRGOrb] = h; DMRGOrb++; } }
         else { for (int h = 0; h < L; h++) if (Ham->getOrbitalIrrep(h) == order[k]) { f1[h] = DMRGOrb; f2[DMRGOrb] = h; DMRGOrb++; } }
      }
      assert(DMRGOrb == L);
   }
};

/* ConvergenceScheme.h:47-98 */
class ConvergenceScheme {
public:
   ConvergenceScheme(const int num_instructions)
      : num(num_instructions), Ds(num_instructions, 0), sweeps(num_instructions, 0), econv(num_instructions, 0.0), noise(num_instructions, 0.0), rtol(num_instructions, 0.0) {}
   virtual ~ConvergenceScheme() {}
   int get_number() const { return num; }
   void set_instruction(const int instruction, const int D, const double energy_conv, const int max_sweeps, const double noise_prefactor, const double davidson_rtol) {
      assert(instruction >= 0 && instruction < num);
      assert(D > 0); assert(energy_conv > 0.0); assert(max_sweeps > 0); assert(davidson_rtol > 0.0);
      Ds[instruction] = D; econv[instruction] = energy_conv; sweeps[instruction] = max_sweeps; noise[instruction] = noise_prefactor; rtol[instruction] = davidson_rtol;
   }
   void setInstruction(const int instruction, const int D, const double energy_conv, const int max_sweeps, const double noise_prefactor) {
      set_instruction(instruction, D, energy_conv, max_sweeps, noise_prefactor, DAVIDSON_DMRG_RTOL);
   }
   int get_D(const int i) const { return Ds[i]; }
   double get_energy_conv(const int i) const { return econv[i]; }
   int get_max_sweeps(const int i) const { return sweeps[i]; }
   double get_noise_prefactor(const int i) const { return noise[i]; }
   double get_dvdson_rtol(const int i) const { return rtol[i]; }
private:
   int num;
   std::vector<int> Ds, sweeps;
   std::vector<double> econv, noise, rtol;
};

/* TwoDM.h:57-136: accessors over the spin-summed arrays A and B (DMRG orbital order) filled by b2_dmrg_calc_2rdm */
class TwoDM {
public:
   TwoDM(const Problem* ProbIn) : Prob(ProbIn), L(ProbIn->gL()), two_rdm_A((size_t)L * L * L * L, 0.0), two_rdm_B((size_t)L * L * L * L, 0.0) {}
   double getTwoDMA_DMRG(const int c1, const int c2, const int c3, const int c4) const {
      if (Irreps::directProd(Prob->gIrrep(c1), Prob->gIrrep(c2)) != Irreps::directProd(Prob->gIrrep(c3), Prob->gIrrep(c4))) return 0.0;
      return two_rdm_A[c1 + (size_t)L * (c2 + (size_t)L * (c3 + (size_t)L * c4))];
   }
   double getTwoDMB_DMRG(const int c1, const int c2, const int c3, const int c4) const {
      if (Irreps::directProd(Prob->gIrrep(c1), Prob->gIrrep(c2)) != Irreps::directProd(Prob->gIrrep(c3), Prob->gIrrep(c4))) return 0.0;
      return two_rdm_B[c1 + (size_t)L * (c2 + (size_t)L * (c3 + (size_t)L * c4))];
   }
   double get1RDM_DMRG(const int c1, const int c2) const {
      if (Prob->gIrrep(c1) != Prob->gIrrep(c2)) return 0.0;
      double value = 0.0;
      for (int o = 0; o < L; o++) value += getTwoDMA_DMRG(c1, o, c2, o);
      return value / (Prob->gN() - 1.0);
   }
   double spin_density_dmrg(const int c1, const int c2) const {
      if (Prob->gIrrep(c1) != Prob->gIrrep(c2) || Prob->gTwoS() <= 0) return 0.0;
      double value = (2 - Prob->gN()) * get1RDM_DMRG(c1, c2);
      for (int o = 0; o < L; o++) value -= getTwoDMA_DMRG(c1, o, o, c2) + getTwoDMB_DMRG(c1, o, o, c2);
      return 1.5 * value / (0.5 * Prob->gTwoS() + 1);
   }
   double getTwoDMA_HAM(const int c1, const int c2, const int c3, const int c4) const { return getTwoDMA_DMRG(h(c1), h(c2), h(c3), h(c4)); }
   double getTwoDMB_HAM(const int c1, const int c2, const int c3, const int c4) const { return getTwoDMB_DMRG(h(c1), h(c2), h(c3), h(c4)); }
   double get1RDM_HAM(const int c1, const int c2) const { return get1RDM_DMRG(h(c1), h(c2)); }
   double spin_density_ham(const int c1, const int c2) const { return spin_density_dmrg(h(c1), h(c2)); }
   double trace() const {
      double val = 0.0;
      for (int a = 0; a < L; a++) for (int b = 0; b < L; b++) val += getTwoDMA_DMRG(a, b, a, b);
      return val;
   }
   double energy() const {
      double val = 0.0;
      for (int a = 0; a < L; a++) for (int b = 0; b < L; b++) for (int c = 0; c < L; c++) for (int d = 0; d < L; d++)
         val += getTwoDMA_DMRG(a, b, c, d) * Prob->gMxElement(a, b, c, d);
      return 0.5 * val + Prob->gEconst();
   }
   double* storage_A() { return two_rdm_A.data(); }
   double* storage_B() { return two_rdm_B.data(); }
private:
   const Problem* Prob;
   int L;
   std::vector<double> two_rdm_A, two_rdm_B;
   int h(int ham) const { return Prob->gReorder() ? Prob->gf1(ham) : ham; }
};

/* Correlations.h:106-193 */
class Correlations {
public:
   Correlations(const Problem* ProbIn, TwoDM* the2DMin) : Prob(ProbIn), the2DM(the2DMin), L(ProbIn->gL()) {
      for (int t = 0; t < 5; t++) table[t].assign((size_t)L * L, 0.0);
   }
   double getCspin_DMRG(const int r, const int c) const { return table[0][r + (size_t)L * c]; }
   double getCdens_DMRG(const int r, const int c) const { return table[1][r + (size_t)L * c]; }
   double getCspinflip_DMRG(const int r, const int c) const { return table[2][r + (size_t)L * c]; }
   double getCdirad_DMRG(const int r, const int c) const { return table[3][r + (size_t)L * c]; }
   double getMutualInformation_DMRG(const int r, const int c) const { return table[4][r + (size_t)L * c]; }
   double getCspin_HAM(const int r, const int c) const { return getCspin_DMRG(h(r), h(c)); }
   double getCdens_HAM(const int r, const int c) const { return getCdens_DMRG(h(r), h(c)); }
   double getCspinflip_HAM(const int r, const int c) const { return getCspinflip_DMRG(h(r), h(c)); }
   double getCdirad_HAM(const int r, const int c) const { return getCdirad_DMRG(h(r), h(c)); }
   double getMutualInformation_HAM(const int r, const int c) const { return getMutualInformation_DMRG(h(r), h(c)); }
   double SingleOrbitalEntropy_DMRG(const int index) const {
      const double val4 = 0.5 * the2DM->getTwoDMA_DMRG(index, index, index, index);
      const double val23 = 0.5 * (the2DM->get1RDM_DMRG(index, index) - the2DM->getTwoDMA_DMRG(index, index, index, index));
      const double val1 = 1.0 - val4 - 2 * val23;
      double entropy = 0.0;
      if (val1 > CORRELATIONS_discardEig) entropy -= val1 * std::log(val1);
      if (val23 > CORRELATIONS_discardEig) entropy -= 2 * val23 * std::log(val23);
      if (val4 > CORRELATIONS_discardEig) entropy -= val4 * std::log(val4);
      return entropy;
   }
   double SingleOrbitalEntropy_HAM(const int index) const { return SingleOrbitalEntropy_DMRG(h(index)); }
   double MutualInformationDistance(const double power) const {
      double Idist = 0.0;
      for (int r = 0; r < L; r++) for (int c = 0; c < L; c++) if (r != c) Idist += table[4][r + (size_t)L * c] * std::pow((double)std::abs(r - c), power);
      return Idist;
   }
   double* storage(int t) { return table[t].data(); }   /* 0 Cspin, 1 Cdens, 2 Cspinflip, 3 Cdirad, 4 MutInfo */
private:
   const Problem* Prob;
   TwoDM* the2DM;
   int L;
   std::vector<double> table[5];
   int h(int ham) const { return Prob->gReorder() ? Prob->gf1(ham) : ham; }
};

/* DMRG.h:93-162.  The object owns one library context on the CUDA device `device` (default: environment variable B2_DEVICE, else 0),
 * the bookkeeper, the MPS and every renormalized operator set; Problem and ConvergenceScheme are borrowed, like in the reference. */
class DMRG {
public:
   DMRG(Problem* Probin, ConvergenceScheme* OptSchemeIn, const bool makechkpt = DMRG_storeMpsOnDisk, const std::string tmpfolder = defaultTMPpath, int* occupancies = NULL, int device = -1)
      : Prob(Probin), OptScheme(OptSchemeIn), ctx(NULL), d(NULL), the2DM(NULL), theCorr(NULL), L(Probin->gL()), nStates(1), maxExc(0), Exc_activated(false),
        makecheckpoints(makechkpt), tempfolder(tmpfolder), ops_ready(false), verbose(true) {
      assert(Prob->checkConsistency());
      Prob->construct_mxelem();
      if (device < 0) { const char* e = std::getenv("B2_DEVICE"); device = e ? std::atoi(e) : 0; }
      detail::check(b2_ctx_create(device, &ctx), "b2_ctx_create");
      std::vector<int> irr(L);
      for (int i = 0; i < L; i++) irr[i] = Prob->gIrrep(i);
      detail::check(b2_problem_set(ctx, L, Prob->gSy(), Prob->gN(), Prob->gTwoS(), Prob->gIrrep(), irr.data(), Prob->mx_table(), Prob->gEconst()), "b2_problem_set");
      if (occupancies != NULL && verbose) std::cout << "chemps2_b200: the ROHF occupation guess is not used; the MPS starts from seeded random blocks" << std::endl;
      setupBookkeeperAndMPS();
      PreSolve();
   }
   virtual ~DMRG() {
      delete theCorr; delete the2DM;
      b2_dmrg_destroy(d);
      b2_ctx_destroy(ctx);
   }
   /* DMRG.cpp:257-266 */
   void PreSolve() {
      /* the reference reads Prob->gMxElement live: callers may have changed the table with Problem::setMxElement (tests/test12.cpp.in) */
      detail::check(b2_problem_update_mx(ctx, Prob->mx_table()), "b2_problem_update_mx");
      detail::check(b2_dmrg_presolve(d), "b2_dmrg_presolve");
      TotalMinEnergy = 1e8;
      ops_ready = true;
   }
   /* DMRG.cpp:268-355: per instruction left + right sweeps until the energy of the last site changes by less than energy_conv */
   double Solve() {
      if (!ops_ready) PreSolve();
      bool change = TotalMinEnergy < 1e8;       /* the very first left sweep keeps the virtual dimensions fixed */
      double Energy = 0.0;
      for (int ins = 0; ins < OptScheme->get_number(); ins++) {
         int nIterations = 0;
         double EnergyPrevious = Energy + 10 * OptScheme->get_energy_conv(ins);
         while (std::fabs(Energy - EnergyPrevious) > OptScheme->get_energy_conv(ins) && nIterations < OptScheme->get_max_sweeps(ins)) {
            EnergyPrevious = Energy;
            Energy = half_sweep(false, change, ins, nIterations);
            change = true;
            Energy = half_sweep(true, change, ins, nIterations);
            if (verbose) std::cout << "***     Energy difference with respect to previous leftright sweep = " << std::fabs(Energy - EnergyPrevious) << std::endl
                                   << "******************************************************************" << std::endl;
            if (makecheckpoints) detail::check(b2_dmrg_save_mps(d, MPSstoragename.c_str(), 0), "b2_dmrg_save_mps");
            nIterations++;
         }
         if (verbose) {
            std::cout << "***  Information on completed instruction " << ins << ":" << std::endl;
            std::cout << "***     The reduced virtual dimension DSU(2)               = " << OptScheme->get_D(ins) << std::endl;
            std::cout << "***     The total number of reduced MPS variables          = " << get_num_mps_var() << std::endl;
            std::cout << "***     Minimum energy encountered during all instructions = " << TotalMinEnergy << std::endl;
            std::cout << "***     Minimum energy encountered during the last sweep   = " << LastMinEnergy << std::endl;
            std::cout << "***     Maximum discarded weight during the last sweep     = " << MaxDiscWeightLastSweep << std::endl;
            std::cout << "******************************************************************" << std::endl;
         }
      }
      return TotalMinEnergy;
   }
   void calc2DMandCorrelations() { calc_rdms_and_correlations(false); }
   /* DMRGtechnics.cpp:40-215 without the 3-RDM (outside the path: aborts when asked for) */
   void calc_rdms_and_correlations(const bool do_3rdm, const bool disk_3rdm = false) {
      (void)disk_3rdm;
      if (do_3rdm) { std::fprintf(stderr, "chemps2_b200: the 3-RDM is outside the accelerated path (DESIGN.md, out of scope)\n"); std::abort(); }
      delete theCorr; delete the2DM;
      the2DM = new TwoDM(Prob);
      theCorr = new Correlations(Prob, the2DM);
      const double t0 = detail::now();
      detail::check(b2_dmrg_calc_2rdm(d, the2DM->storage_A(), the2DM->storage_B()), "b2_dmrg_calc_2rdm");
      detail::check(b2_dmrg_calc_correlations(d, the2DM->storage_A(), the2DM->storage_B(), theCorr->storage(0), theCorr->storage(1), theCorr->storage(2),
                                              theCorr->storage(3), theCorr->storage(4)), "b2_dmrg_calc_correlations");
      ops_ready = false;       /* the chain now holds the reduced operator sets of the 2-RDM sweep */
      if (verbose) {
         std::cout << "******************************************************************" << std::endl;
         std::cout << "***  Information on the 2-RDM and correlation sweeps:" << std::endl;
         std::cout << "***     Elapsed wall time        = " << detail::now() - t0 << " seconds" << std::endl;
         std::cout << "***     2-RDM trace              = " << the2DM->trace() << " (should be " << Prob->gN() * (Prob->gN() - 1.0) << ")" << std::endl;
         std::cout << "***     2-RDM energy             = " << the2DM->energy() << std::endl;
         std::cout << "******************************************************************" << std::endl;
      }
   }
   TwoDM* get2DM() { return the2DM; }
   Correlations* getCorrelations() { return theCorr; }
   void deleteStoredMPS() { if (makecheckpoints) std::remove(MPSstoragename.c_str()); }
   void deleteStoredOperators() {}          /* operators live in HBM / pinned host memory and die with the object */
   /* DMRG.cpp:464-505 */
   void activateExcitations(const int maxExcIn) { Exc_activated = true; maxExc = maxExcIn; }
   void newExcitation(const double EshiftIn) {
      assert(Exc_activated);
      assert(nStates - 1 < maxExc);
      delete theCorr; theCorr = NULL; delete the2DM; the2DM = NULL;
      nStates++;
      set_storage_name();
      detail::check(b2_dmrg_new_excitation(d, EshiftIn, OptScheme->get_D(0), Initialize::seed() + 7919ULL * (nStates - 1)), "b2_dmrg_new_excitation");
      PreSolve();
   }
   int get_num_mps_var() const {
      long long n = 0;
      for (int site = 0; site < L; site++) n += b2_dmrg_mps_size(d, site);
      return (int)n;
   }
   static void PrintLicense() {}
   /* public members of the reference's DMRG that are not on the accelerated path (FCI coefficients of the MPS: DMRGtechnics.cpp:217-507;
    * 4-RDM contraction for CASPT2: DMRGfock.cpp): callers compile, calling them aborts with a message instead of returning wrong numbers */
   double getSpecificCoefficient(int*) const { not_built("DMRG::getSpecificCoefficient"); return 0.0; }
   double getFCIcoefficient(int*, int*, const bool = true) const { not_built("DMRG::getFCIcoefficient"); return 0.0; }
   void Symm4RDM(double*, const int, const int, const bool) { not_built("DMRG::Symm4RDM"); }
   /* ---- beyond the reference's interface ---- */
   void set_verbose(bool v) { verbose = v; }
   /* one process per GPU: shard sigma terms and operator updates over `world` ranks; fn sums a device vector over them (NCCL) */
   void set_world(int world, int rank, b2_allreduce_fn fn, void* user) { detail::check(b2_dmrg_set_world(d, world, rank, fn, user), "b2_dmrg_set_world"); PreSolve(); }
   /* keep only the operator sets in use in HBM, the rest in pinned host memory (the reference's disk mode) */
   void set_spill(bool on) { detail::check(b2_dmrg_set_spill(d, on ? 1 : 0), "b2_dmrg_set_spill"); }
   b2_dmrg* handle() { return d; }
   b2_ctx* context() { return ctx; }
private:
   Problem* Prob;
   ConvergenceScheme* OptScheme;
   b2_ctx* ctx;
   b2_dmrg* d;
   TwoDM* the2DM;
   Correlations* theCorr;
   int L, nStates, maxExc;
   bool Exc_activated, makecheckpoints;
   std::string tempfolder, MPSstoragename;
   bool ops_ready, verbose;
   double TotalMinEnergy = 1e8, LastMinEnergy = 1e8, MaxDiscWeightLastSweep = 0.0;

   static void not_built(const char* what) {
      std::fprintf(stderr, "chemps2_b200: %s is outside the accelerated two-site sweep path (DESIGN.md, out of scope)\n", what);
      std::abort();
   }
   void set_storage_name() {
      std::stringstream s;
      s << DMRG_MPS_storage_prefix << nStates - 1 << ".b2mps";
      MPSstoragename = s.str();
   }
   /* DMRG.cpp:123-207: bookkeeper at the first instruction's D, then either the checkpoint or a random left-normalised MPS */
   void setupBookkeeperAndMPS() {
      detail::check(b2_bk_init(ctx, OptScheme->get_D(0)), "b2_bk_init");
      detail::check(b2_dmrg_create(ctx, &d), "b2_dmrg_create");
      set_storage_name();
      struct stat info;
      const bool loadedMPS = makecheckpoints && stat(MPSstoragename.c_str(), &info) == 0;
      if (loadedMPS) {
         int converged = 0;
         detail::check(b2_dmrg_load_mps(d, MPSstoragename.c_str(), &converged), "b2_dmrg_load_mps");
         if (verbose) std::cout << "Loaded MPS " << MPSstoragename << " converged y/n? : " << converged << std::endl;
      } else {
         detail::check(b2_dmrg_random_mps(d, Initialize::seed()), "b2_dmrg_random_mps");
      }
   }
   double half_sweep(bool to_right, bool change, int ins, int nIterations) {
      double tm[5], emin = 0.0, dw = 0.0, info[4];
      b2_dmrg_timers(d, tm, 1);
      const double t0 = detail::now();
      detail::check(b2_dmrg_sweep(d, to_right ? 1 : 0, OptScheme->get_dvdson_rtol(ins), OptScheme->get_noise_prefactor(ins), OptScheme->get_D(ins), change ? 1 : 0, &emin, &dw), "b2_dmrg_sweep");
      const double elapsed = detail::now() - t0;
      b2_dmrg_timers(d, tm, 0);
      detail::check(b2_dmrg_sweep_info(d, info), "b2_dmrg_sweep_info");
      LastMinEnergy = info[1]; MaxDiscWeightLastSweep = info[2];
      if (LastMinEnergy < TotalMinEnergy) TotalMinEnergy = LastMinEnergy;
      if (verbose) {
         std::cout << "******************************************************************" << std::endl;
         std::cout << "***  Information on " << (to_right ? "right" : "left") << " sweep " << nIterations << " of instruction " << ins << ":" << std::endl;
         std::cout << "***     Elapsed wall time        = " << elapsed << " seconds" << std::endl;
         std::cout << "***       |--> plan building     = " << tm[0] << " seconds" << std::endl;
         std::cout << "***       |--> S.solve           = " << tm[1] << " seconds (" << (long long)tm[4] << " sigma builds)" << std::endl;
         std::cout << "***       |--> S.split           = " << tm[2] << " seconds" << std::endl;
         std::cout << "***       |--> Tensor update     = " << tm[3] << " seconds" << std::endl;
         std::cout << "***     Minimum energy           = " << LastMinEnergy << std::endl;
         std::cout << "***     Maximum discarded weight = " << MaxDiscWeightLastSweep << std::endl;
         if (!to_right) std::cout << "******************************************************************" << std::endl;
      }
      return info[0];
   }
};

}   /* namespace */

#endif
