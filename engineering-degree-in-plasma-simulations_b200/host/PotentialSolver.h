// Facade of ch4/v3/src/PotentialSolver.h.  solve()/solveGS() run the red-black SOR kernels (same node classes,
// update formula, residual rule and SOR weight as PotentialSolver.cpp:69-166); computeEF() the gradient kernel.
// solve() with SolverType::PCG runs the GS path with the reference's GS iteration budget (20 x max_it, PotentialSolver.cpp:46-49):
// the reference's PCG matrix swaps the x/z coefficients on non-cubic cells, is not symmetric, and its CG diverges on
// the discharge meshes (its solveGSlinear fallback does the work), so the default keeps the equation solveGS relaxes.
// solveNRPCG() is the device port of PotentialSolver::solveNRPCG with the reference's matrix, for callers that want its results.
#ifndef POTENTIALSOLVER_H
#define POTENTIALSOLVER_H
#include <istream>
#include <memory>
#include <ostream>
#include "World.h"
#include "picgpu.h"

enum SolverType { GS, PCG, QN };

class PotentialSolver {
protected:
    World& world;
    const SolverType solver_type;
    unsigned PCG_max_solver_it, GS_max_solver_it;
    type_calc tolerance;
    std::shared_ptr<picg_solver_s> handle;
    unsigned last_iterations = 0;
    type_calc last_L2 = 0;

public:
    PotentialSolver(World& world, unsigned max_solver_it, type_calc tolerance, SolverType solver_type);
    bool solve();
    bool solveGS();
    bool solveNRPCG();
    void computeEF();
    void setReferenceValues(type_calc phi0, type_calc n0, type_calc Te0);
    unsigned get_GS_max_it() { return GS_max_solver_it; }
    unsigned get_PCG_max_it() { return PCG_max_solver_it; }
    unsigned iterations() const { return last_iterations; }
    type_calc residual() const { return last_L2; }
};
std::ostream& operator<<(std::ostream& out, SolverType& type);
std::istream& operator>>(std::istream& in, SolverType& type);
#endif
