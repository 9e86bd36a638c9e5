// Facade of ch4/v3/src/Species.h.  The particle store is a structure of arrays on the GPU; this object holds a
// ref-counted handle to it, so std::vector<Species> may relocate elements freely (ch4/v3/src/main.cpp:71,105-108).
#ifndef SPECIES_H
#define SPECIES_H
#include <memory>
#include <string>
#include <vector>
#include "Field.h"
#include "Rnd.h"
#include "Vec3.h"
#include "World.h"
#include "all.h"
#include "picgpu.h"

class Particle {
public:
    type_calc3 pos, vel;
    type_calc macro_weight;
    Particle(type_calc x, type_calc y, type_calc z, type_calc u, type_calc v, type_calc w, type_calc mpw) noexcept : pos{x, y, z}, vel{u, v, w}, macro_weight{mpw} {}
    Particle(type_calc3 p, type_calc3 v, type_calc mpw) noexcept : pos{p}, vel{v}, macro_weight{mpw} {}
};

class Species {
protected:
    World& world;
    std::shared_ptr<picg_species_s> handle;
    std::vector<Particle> particles_mirror;           // refreshed by getPartRef()/getConstPartRef()
    bool sorted = false;
    void bindFields();

public:
    const std::string name;
    const type_calc mass, charge, mpw0;
    const type_calc E_ion = -666;

    Field<type_calc> den, den_avg, T;
    Field<type_calc3> vel;
    Field<type_calc> macro_part_count;

    Species(std::string name, type_calc mass, type_calc charge, World& world, type_calc mpw0);
    Species(std::string name, type_calc mass, type_calc charge, World& world, type_calc mpw0, type_calc E_ion);
    Species(const Species& o);                         // shares the device store (vector relocation)

    picg_species_t dev() const { return handle.get(); }
    size_t getNumParticles() const;
    type_calc advance_time_multi = -1, advance_time_serial = 0;
    void advanceNonElectron(Species& neutrals, Species& spherium, type_calc dt);
    void advanceElectrons(type_calc dt);
    void computeNumberDensity();
    void addParticle(type_calc x, type_calc y, type_calc z, type_calc u, type_calc v, type_calc w, type_calc macro_weight);
    void addParticle(type_calc3 pos, type_calc3 vel, type_calc macro_weight);
    void addParticle(type_calc3 pos, type_calc3 vel) { addParticle(pos, vel, mpw0); }
    void loadParticleBoxThermal(type_calc3 x0, type_calc3 sides, type_calc num_den, type_calc T);
    bool setSorted(bool s) noexcept { return sorted = s; }
    bool isSorted() noexcept { return sorted; }

    type_calc getMicroCount();
    type_calc3 getMomentum();
    type_calc getKE();
    void updateAverages();
    void sampleMoments();
    void computeGasProperties();
    void clearSamples();
    void computeMacroParticlesCount();
    const std::vector<Particle>& getConstPartRef();
    const Particle& getConstPartRef(int i);
    void setParticles(const std::vector<Particle>& p);  // uploads a host particle set (raw store, no addParticle filtering)
    void merge();                                        // Species.cpp:1037-1145 -> picg_species_merge
    void sortByCell();
};
#endif
