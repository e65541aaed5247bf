/* A caller written against the C++ mirror of the reference's public classes (include/chemps2_b200.hpp), in the way the reference's
 * own tests drive a calculation (tests/test1.cpp.in, test2.cpp.in, test5.cpp.in): Hamiltonian from an FCIDUMP -> Problem ->
 * ConvergenceScheme -> DMRG::Solve -> calc2DMandCorrelations -> excited states.  tests/test_cpp_mirror.py writes the FCIDUMP from
 * the committed fixtures, runs this binary and checks the numbers it prints.
 *
 *   dmrg_caller host  <fcidump> <group> <TwoS> <N> <Irrep> <none|d2h|c2v> <out.bin>     host classes only (no GPU): dumps irreps + folded table
 *   dmrg_caller accessors <fcidump> <group> <TwoS> <N> <Irrep> <none|d2h|c2v> <out.bin>  TwoDM / Correlations accessor arithmetic on filled-in arrays (no GPU)
 *   dmrg_caller pairing                                                                    the reference's tests/test12.cpp.in on the GPU: reduced BCS model, L = 8
 *   dmrg_caller solve <fcidump> <group> <TwoS> <N> <Irrep> <none|d2h|c2v> <D> <n_excited>  whole calculation on the GPU, one JSON line */
#include <cstring>

#include "chemps2_b200.hpp"

using std::cout;
using std::endl;

static void reorder(CheMPS2::Problem* Prob, const char* how) {
   if (!std::strcmp(how, "d2h")) Prob->SetupReorderD2h();
   if (!std::strcmp(how, "c2v")) Prob->SetupReorderC2v();
}

/* tests/test12.cpp.in:36-103 — the model Hamiltonian is written into the folded table with Problem::setMxElement AFTER the DMRG object
 * exists, PreSolve rebuilds the operators, then Solve with the two-instruction scheme of the test; known answer -25.5134137600604 */
static int pairing_model() {
   CheMPS2::Initialize::Init();
   const int L = 8;
   double eps[] = {-3.5, -2.5, -1.5, -0.5, 0.5, 1.5, 2.5, 3.5};
   const double g = -1.0, power = 0.0;
   const int N = L, TwoS = 0, Irrep = 0;
   CheMPS2::ConvergenceScheme* OptScheme = new CheMPS2::ConvergenceScheme(2);
   OptScheme->setInstruction(0, 100, 1e-10, 10, 0.5);
   OptScheme->setInstruction(1, 1000, 1e-10, 10, 0.0);
   const int group = 0;
   int* irreps = new int[L];
   for (int orb = 0; orb < L; orb++) irreps[orb] = 0;
   CheMPS2::Hamiltonian* Ham = new CheMPS2::Hamiltonian(L, group, irreps);
   delete[] irreps;
   CheMPS2::Problem* Prob = new CheMPS2::Problem(Ham, TwoS, N, Irrep);
   CheMPS2::DMRG* theDMRG = new CheMPS2::DMRG(Prob, OptScheme);
   for (int orb1 = 0; orb1 < L; orb1++)
      for (int orb2 = 0; orb2 < L; orb2++) {
         const double eri = g * std::pow(std::fabs(eps[orb1] * eps[orb2]), power);
         const double oei = (eps[orb1] + eps[orb2]) / (N - 1);
         if (orb1 == orb2) Prob->setMxElement(orb1, orb1, orb2, orb2, eri + oei);
         else { Prob->setMxElement(orb1, orb1, orb2, orb2, eri); Prob->setMxElement(orb1, orb2, orb1, orb2, oei); }
      }
   theDMRG->PreSolve();
   const double Energy = theDMRG->Solve();
   theDMRG->calc2DMandCorrelations();
   const double rdm_energy = theDMRG->get2DM()->energy(), trace = theDMRG->get2DM()->trace();
   double pair_occupation = 0.0;                          /* seniority-zero state: <n_i n_i> pairs only, A(i,i,i,i) = 2 <n_up n_down> */
   for (int i = 0; i < L; i++) pair_occupation += 0.5 * theDMRG->get2DM()->getTwoDMA_HAM(i, i, i, i);
   delete theDMRG; delete OptScheme; delete Prob; delete Ham;
   std::printf("B2JSON {\"energy\": %.12f, \"rdm_energy\": %.12f, \"trace\": %.10f, \"pairs\": %.10f}\n", Energy, rdm_energy, trace, pair_occupation);
   return std::fabs(Energy + 25.5134137600604) < 1e-8 ? 0 : 7;
}

int main(int argc, char** argv) {
   if (argc > 1 && std::string(argv[1]) == "pairing") return pairing_model();
   if (argc < 9) { std::fprintf(stderr, "usage: see the header of tests/cpp/dmrg_caller.cpp\n"); return 2; }
   const std::string mode = argv[1], matrixelements = argv[2];
   const int psi4groupnumber = std::atoi(argv[3]), TwoS = std::atoi(argv[4]), N = std::atoi(argv[5]), Irrep = std::atoi(argv[6]);

   CheMPS2::Initialize::Init();
   CheMPS2::Hamiltonian* Ham = new CheMPS2::Hamiltonian(matrixelements, psi4groupnumber);
   CheMPS2::Problem* Prob = new CheMPS2::Problem(Ham, TwoS, N, Irrep);
   reorder(Prob, argv[7]);
   const int L = Prob->gL();

   if (mode == "host") {
      Prob->construct_mxelem();
      FILE* f = std::fopen(argv[8], "wb");
      if (!f) return 3;
      std::fwrite(&L, sizeof(int), 1, f);
      for (int i = 0; i < L; i++) { const int ir = Prob->gIrrep(i); std::fwrite(&ir, sizeof(int), 1, f); }
      const double econst = Prob->gEconst();
      std::fwrite(&econst, sizeof(double), 1, f);
      std::fwrite(Prob->mx_table(), sizeof(double), (size_t)L * L * L * L, f);
      std::fclose(f);
      /* write -> read round trip of the FCIDUMP writer */
      const std::string copy = std::string(argv[8]) + ".fcidump";
      Ham->writeFCIDUMP(copy, N, TwoS, Irrep);
      CheMPS2::Hamiltonian again(copy, psi4groupnumber);
      double worst = std::fabs(again.getEconst() - Ham->getEconst());
      for (int a = 0; a < L; a++) for (int b = 0; b < L; b++) {
         worst = std::max(worst, std::fabs(again.getTmat(a, b) - Ham->getTmat(a, b)));
         for (int c = 0; c < L; c++) for (int d = 0; d < L; d++) worst = std::max(worst, std::fabs(again.getVmat(a, b, c, d) - Ham->getVmat(a, b, c, d)));
      }
      std::remove(copy.c_str());
      std::printf("{\"L\": %d, \"reorder\": %d, \"fcidump_roundtrip\": %.3e}\n", L, Prob->gReorder() ? 1 : 0, worst);
      delete Prob; delete Ham;
      return 0;
   }

   if (mode == "accessors") {
      /* A, B and the correlation tables filled with a fixed pattern; every accessor evaluated in Hamiltonian AND DMRG orbital order */
      Prob->construct_mxelem();
      CheMPS2::TwoDM dm(Prob);
      CheMPS2::Correlations corr(Prob, &dm);
      const size_t L4 = (size_t)L * L * L * L;
      for (size_t i = 0; i < L4; i++) { dm.storage_A()[i] = std::sin(0.37 * (double)i + 0.1); dm.storage_B()[i] = std::cos(0.23 * (double)i - 0.4); }
      for (int t = 0; t < 5; t++) for (int i = 0; i < L * L; i++) corr.storage(t)[i] = std::sin(1.0 + t + 0.61 * i);
      std::vector<double> out;
      out.push_back(dm.trace()); out.push_back(dm.energy()); out.push_back(corr.MutualInformationDistance(2.0));
      for (int i = 0; i < L; i++) { out.push_back(corr.SingleOrbitalEntropy_HAM(i)); out.push_back(corr.SingleOrbitalEntropy_DMRG(i)); }
      for (int i = 0; i < L; i++) for (int j = 0; j < L; j++) {
         out.push_back(dm.get1RDM_HAM(i, j)); out.push_back(dm.get1RDM_DMRG(i, j));
         out.push_back(dm.spin_density_ham(i, j)); out.push_back(dm.spin_density_dmrg(i, j));
         out.push_back(corr.getCspin_HAM(i, j)); out.push_back(corr.getCdens_HAM(i, j)); out.push_back(corr.getCspinflip_HAM(i, j));
         out.push_back(corr.getCdirad_HAM(i, j)); out.push_back(corr.getMutualInformation_HAM(i, j)); out.push_back(corr.getMutualInformation_DMRG(i, j));
         out.push_back(dm.getTwoDMA_HAM(i, j, (i + 1) % L, (j + 2) % L)); out.push_back(dm.getTwoDMB_HAM(i, j, j, i));
      }
      FILE* f = std::fopen(argv[8], "wb");
      if (!f) return 3;
      std::fwrite(out.data(), sizeof(double), out.size(), f);
      std::fclose(f);
      std::printf("{\"L\": %d, \"values\": %zu}\n", L, out.size());
      delete Prob; delete Ham;
      return 0;
   }

   const int D = std::atoi(argv[8]);
   const int n_excited = argc > 9 ? std::atoi(argv[9]) : 0;
   CheMPS2::ConvergenceScheme* OptScheme = new CheMPS2::ConvergenceScheme(1);
   OptScheme->setInstruction(0, D, 1e-10, 30, 0.0);

   CheMPS2::DMRG* theDMRG = new CheMPS2::DMRG(Prob, OptScheme);
   std::vector<double> energies, rdm_energies, traces;
   energies.push_back(theDMRG->Solve());
   theDMRG->calc2DMandCorrelations();
   rdm_energies.push_back(theDMRG->get2DM()->energy());
   traces.push_back(theDMRG->get2DM()->trace());
   /* observables of the ground state in Hamiltonian orbital order */
   double n_elec = 0.0, entropy = 0.0, mutinfo = 0.0, worst_1rdm_asym = 0.0;
   for (int i = 0; i < L; i++) {
      n_elec += theDMRG->get2DM()->get1RDM_HAM(i, i);
      entropy += theDMRG->getCorrelations()->SingleOrbitalEntropy_HAM(i);
      for (int j = 0; j < L; j++) {
         mutinfo += theDMRG->getCorrelations()->getMutualInformation_HAM(i, j);
         worst_1rdm_asym = std::max(worst_1rdm_asym, std::fabs(theDMRG->get2DM()->get1RDM_HAM(i, j) - theDMRG->get2DM()->get1RDM_HAM(j, i)));
      }
   }
   if (n_excited > 0) theDMRG->activateExcitations(n_excited);
   for (int s = 0; s < n_excited; s++) {
      theDMRG->newExcitation(20.0);
      energies.push_back(theDMRG->Solve());
      theDMRG->calc2DMandCorrelations();
      rdm_energies.push_back(theDMRG->get2DM()->energy());
      traces.push_back(theDMRG->get2DM()->trace());
   }
   if (CheMPS2::DMRG_storeMpsOnDisk) theDMRG->deleteStoredMPS();
   if (CheMPS2::DMRG_storeRenormOptrOnDisk) theDMRG->deleteStoredOperators();
   delete theDMRG; delete OptScheme; delete Prob; delete Ham;

   std::printf("B2JSON {\"energies\": [");
   for (size_t i = 0; i < energies.size(); i++) std::printf("%s%.12f", i ? ", " : "", energies[i]);
   std::printf("], \"rdm_energies\": [");
   for (size_t i = 0; i < rdm_energies.size(); i++) std::printf("%s%.12f", i ? ", " : "", rdm_energies[i]);
   std::printf("], \"traces\": [");
   for (size_t i = 0; i < traces.size(); i++) std::printf("%s%.10f", i ? ", " : "", traces[i]);
   std::printf("], \"n_elec\": %.10f, \"entropy_sum\": %.10f, \"mutinfo_sum\": %.10f, \"rdm1_asym\": %.3e}\n", n_elec, entropy, mutinfo, worst_1rdm_asym);
   return 0;
}
