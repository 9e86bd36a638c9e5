// Facade of ch4/v3/src/World.h.  Same public surface (constructors, geometry helpers, time stepping, objects,
// the public Field members phi/rho/node_vol/ef/object_id/node_type); the fields live on the GPU behind a
// ref-counted handle and are mirrored lazily (Field.h).  Every method forwards to the C ABI (include/picgpu.h).
#ifndef WORLD_H
#define WORLD_H
#include <chrono>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "Field.h"
#include "Object.h"
#include "Vec3.h"
#include "all.h"
#include "picgpu.h"

class Species;

class World {
protected:
    enum NodeType { REGULAR, NEUMANN, DIRICHLET };
    type_calc3 x0, dx, inv_dx, xm, xc;
    type_calc dt = 1e-4;
    int num_ts = 0, ts = -1;
    type_calc time = 0;
    std::chrono::time_point<std::chrono::high_resolution_clock> time_start;
    bool steady_state = false;
    std::vector<std::unique_ptr<Object>> objects;
    std::shared_ptr<picg_world_s> handle;          // device resources (ref-counted: handles outlive relocations)
    void bindFields();
    void registerObject(Object* o);

public:
    const int nn[3];
    const int ni, nj, nk;
    const int ni_1, nj_1, nk_1;
    const int nv;
    const int num_cells;

    Field<type_calc> phi, rho, node_vol;
    Field<type_calc3> ef;
    Field<int> object_id;
    Field<type_calc> object_phi;
    Field<int> node_type;

    World(int ni, int nj, int nk, type_calc x1, type_calc y1, type_calc z1, type_calc x2, type_calc y2, type_calc z2);
    World(int ni, int nj, int nk, type_calc3 vec1, type_calc3 vec2);
    World(const World&) = delete;

    picg_world_t dev() const { return handle.get(); }
    void syncToDevice();                            // uploads host-modified mirrors (phi, rho, ef)
    void deviceChanged(int field);                  // marks a mirror stale

    type_calc3 getX0() const { return x0; }
    type_calc3 getDx() const { return dx; }
    type_calc3 getXm() const { return xm; }
    type_calc3 getXc() const { return xc; }
    type_calc3 getL() const { return {dx[0] * ni_1, dx[1] * nj_1, dx[2] * nk_1}; }     // World.cpp:94-100
    type_calc getCellVolume() const { return dx[0] * dx[1] * dx[2]; }
    int getNumCells() const { return num_cells; }
    type_calc getPE();
    type_calc3 XtoL(const type_calc3& x) const { return (x - x0).elWiseMult(inv_dx); }  // World.cpp:123-127
    int3 XtoIJK(const type_calc3& x) const { type_calc3 l = XtoL(x); return {int(l[0]), int(l[1]), int(l[2])}; }
    int XtoC(const type_calc3& x) const { type_calc3 l = XtoL(x); return (int(l[2]) * nj_1 + int(l[1])) * ni_1 + int(l[0]); }
    type_calc3 LtoX(int i, int j, int k) const { return {x0[0] + i * dx[0], x0[1] + j * dx[1], x0[2] + k * dx[2]}; }
    bool steadyState() const { return steady_state; }

    void computeChargeDensity(std::vector<Species>& species);
    bool inBounds(const type_calc3& pos) const;
    void addInlet(std::string face, type_calc phi_set = 0, int node_type = DIRICHLET);

    void computeObjectID();
    int inObject(const type_calc3& pos) const;
    template <class T, class... Args> void addObject(Args&&... args);
    std::string printObjects() const;

    void setTime(type_calc dt, int num_ts);
    type_calc getDt() const { return dt; }
    int getTs() const { return ts; }
    type_calc getTime() const { return time; }
    bool isLastTimeStep() const { return ts == num_ts - 1; }        // World.cpp:334-336
    bool advanceTime() { time += dt; ts++; return ts <= num_ts; }   // World.cpp:337-341 (num_ts+1 iterations)
    type_calc getWallTime();
    void setTimeStart() { time_start = std::chrono::high_resolution_clock::now(); }
};

template <class T, class... Args>
void World::addObject(Args&&... args) {
    static_assert(std::is_base_of<Object, T>::value, "T must derive from Object");
    try {
        objects.emplace_back(std::make_unique<T>(std::forward<Args>(args)...));
        registerObject(objects.back().get());
    } catch (const std::invalid_argument& e) {                       // World.h:121-128: constructor errors are reported, not propagated
        std::cerr << "Error adding object: " << e.what() << '\n';
    } catch (...) {
        std::cerr << "Unknown error occurred while adding object\n";
    }
}
int inletName2Index(std::string face_name);
#endif
