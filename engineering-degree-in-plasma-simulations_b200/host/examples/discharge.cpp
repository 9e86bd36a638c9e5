// Client of the facade written the way ch4/v3/src/main.cpp drives the reference (same calls, same order,
// main.cpp:63-288), with the Poisson solve live in the loop as in ch2/ch3/ch4-v1.  Used by tests/test_facade.py:
// the per-step diagnostics printed here must equal those of the same scenario driven through the ctypes binding.
#include <iostream>
#include <memory>
#include "Config.h"
#include "Interactions.h"
#include "Outputs.h"
#include "PotentialSolver.h"
#include "Source.h"
#include "Species.h"
#include "World.h"
#include "funkc.h"

int main(int argc, char* argv[]) {
    std::vector<std::string> args(argv + 1, argv + argc);
    Config& config = Config::getInstance();
    config.setSUBCYCLING(parseArgument(args, "--subcycling", false));
    config.setMERGING(parseArgument(args, "--merging", false));
    int num_ts = parseArgument(args, "--num_ts", 10);
    type_calc dt = parseArgument(args, "--dt", 1e-12);
    type_calc phi = parseArgument(args, "--phi", -4000.0);
    int solver_max_it = parseArgument(args, "--s_max_it", 200);
    type_calc solver_tol = parseArgument(args, "--s_tol", 1.0);
    unsigned seed = parseArgument(args, "--seed", 1234u);
    int n_ele = parseArgument(args, "--electrons", 20000);
    std::string table = parseArgument(args, "--table", std::string("data/Oxygen_momentum_transfer.txt"));
    bool extras = parseArgument(args, "--extras", false);      // after the loop: DSMC_MEX, binary field output, checkpoint round trip, NR-PCG
    rnd = Rnd(seed);

    std::unique_ptr<World> world = std::make_unique<World>(21, 21, 31, type_calc3{-0.004, -0.004, 0.0}, type_calc3{0.004, 0.004, 0.005});
    world->setTimeStart();
    world->setTime(dt, num_ts);
    world->addObject<Rectangle>(type_calc3(world->getXc()[0], world->getXc()[1], world->getX0()[2]), phi, type_calc3(world->getL()[0], world->getL()[1], world->getL()[2] * 0.1));
    world->addObject<Rectangle>(type_calc3(world->getXc()[0], world->getXc()[1], world->getXm()[2]), -phi, type_calc3(world->getL()[0], world->getL()[1], world->getL()[2] * 0.1));
    world->computeObjectID();

    type_calc E_ion_O = 1313.9 * 1000 / Const::N_a;
    std::vector<Species> species;
    species.emplace_back("O", 16 * Const::amu, 0, *world, 5e11, E_ion_O);
    species.emplace_back("O+", 16 * Const::amu, Const::q_e, *world, 100);
    species.emplace_back("e-", Const::m_e, -Const::q_e, *world, 100);
    type_calc3 L = world->getL();
    type_calc gap_vol = L[0] * L[1] * L[2] * 0.8;
    species[0].loadParticleBoxThermal(world->getXc(), type_calc3(L[0], L[1], L[2] * 0.8), 2e5 * 5e11 / gap_vol, 300);
    species[2].loadParticleBoxThermal(world->getXc(), type_calc3(L[0], L[1], L[2] * 0.8), n_ele * 100.0 / gap_vol, 3000);

    std::vector<std::unique_ptr<Interaction>> interactions;
    try {
        interactions.emplace_back(std::make_unique<MC_MEX_Ionization>(species[0], species[1], species[2], *world, table));
    } catch (const std::invalid_argument& e) { std::cerr << "no collisions: " << e.what() << "\n"; }
    std::vector<std::unique_ptr<Source>> sources;
    PotentialSolver solver(*world, solver_max_it, solver_tol, GS);
    solver.setReferenceValues(0, 0, 1e20);

    Species& neutral_oxygen = species[0];
    Species& electrons = species[2];
    for (Species& sp : species) sp.computeMacroParticlesCount();
    solver.solve();
    solver.computeEF();
    while (world->advanceTime()) {
        for (auto& s : sources) s->sample();
        for (auto& i : interactions) i->apply(world->getDt());
        for (Species& sp : species) {
            if (sp.name == electrons.name) sp.advanceElectrons(dt);
            else sp.advanceNonElectron(neutral_oxygen, neutral_oxygen, dt);
            sp.computeNumberDensity();
            sp.sampleMoments();
            sp.computeMacroParticlesCount();
        }
        if (world->getTs() > 5) for (Species& sp : species) sp.updateAverages();
        world->computeChargeDensity(species);
        solver.solve();
        solver.computeEF();
        std::cout << "STEP " << world->getTs();
        for (Species& sp : species) std::cout << " " << sp.name << " " << sp.getNumParticles() << " " << std::setprecision(17) << sp.getKE();
        std::cout << " PE " << world->getPE() << " phi_mid " << world->phi[10][10][15] << " rho_mid " << world->rho[10][10][15] << " it " << solver.iterations() << "\n";
    }
    if (extras) {
        // neutral-neutral collisions (the ch4/v2-SPHERE interaction, ch4/v2/main.cpp:266)
        DSMC_MEX dsmc(neutral_oxygen, *world);
        type_calc ke0 = neutral_oxygen.getKE();
        dsmc.apply(1e-6);
        std::cout << "EXTRA dsmc candidates " << dsmc.stats().candidates << " collisions " << dsmc.stats().collisions << " ke_ratio " << std::setprecision(17)
                  << neutral_oxygen.getKE() / ke0 << "\n";
        // binary .vti (needs results/, like the reference) and a checkpoint round trip
        Output::fieldsOutput(*world, species, "extra_");
        std::vector<size_t> before; for (Species& sp : species) before.push_back(sp.getNumParticles());
        type_calc phi_mid = world->phi[10][10][15];
        bool ok = Output::saveCheckpoint("results/run.ckp", *world, species);
        electrons.advanceElectrons(dt);                         // change the state, then restore it
        world->phi = 0;
        ok = ok && Output::loadCheckpoint("results/run.ckp", *world, species);
        bool same = true; for (size_t k = 0; k < species.size(); k++) same = same && species[k].getNumParticles() == before[k];
        std::cout << "EXTRA checkpoint ok " << ok << " counts_restored " << same << " phi_restored " << (world->phi[10][10][15] == phi_mid) << "\n";
        // the reference's other solver
        PotentialSolver pcg(*world, 500, 1e-4, PCG);
        pcg.setReferenceValues(0, 0, 1e20);
        std::cout << "EXTRA nrpcg converged " << pcg.solveNRPCG() << " phi_mid " << world->phi[10][10][15] << "\n";
    }
    std::cout << "Simulation took " << world->getWallTime() << " seconds." << std::endl;
    return 0;
}
