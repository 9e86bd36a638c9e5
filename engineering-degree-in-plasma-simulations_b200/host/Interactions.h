// Facade of ch4/v3/src/Interactions.h: the Interaction interface, MC_MEX_Ionization (the one interaction the v3 main
// loop instantiates, main.cpp:122; per-cell Monte-Carlo kernel csrc/mcc.cu) and DSMC_MEX (the neutral-neutral NTC
// collisions of the ch4/v1 and ch4/v2-SPHERE mains, ch4/v2/main.cpp:266; csrc/dsmc.cu).
#ifndef INTERACTIONS_H
#define INTERACTIONS_H
#include <memory>
#include <stdexcept>
#include <string>
#include "Species.h"
#include "World.h"
#include "picgpu.h"

class Interaction {
public:
    virtual void apply(type_calc dt) noexcept = 0;
    virtual ~Interaction() noexcept = default;
};

class DSMC_MEX : public Interaction {
protected:
    Species &species1, &species2;
    World& world;
    std::shared_ptr<picg_dsmc_s> handle;
    picg_dsmc_stats last{};

public:
    DSMC_MEX(Species& species, World& world) noexcept;                     // collisions within one species (Interactions.cpp:143-157)
    DSMC_MEX(Species& species1, Species& species2, World& world);          // throws std::invalid_argument when the mpw0 differ (:158-176)
    void apply(type_calc dt) noexcept override;
    const picg_dsmc_stats& stats() const { return last; }
};

class MC_MEX_Ionization : public Interaction {
protected:
    Species &neutrals, &ions, &electrons;
    World& world;
    std::shared_ptr<picg_mcc_s> handle;
    picg_mcc_stats last{};

public:
    // throws std::invalid_argument like the reference (Interactions.cpp:479-489,521): bad weights, no E_ion, unreadable table
    MC_MEX_Ionization(Species& neutrals, Species& ions, Species& electrons, World& world,
                      std::string coll_data_path = "data/Oxygen_momentum_transfer.txt", int freq_should_use_map = 10);
    void apply(type_calc dt) noexcept override;
    const picg_mcc_stats& stats() const { return last; }
    // The same class in ch4/v2 (ch4/v2/Interactions.cpp:476-735) uses FIXED weights: Bird's candidate count with neutrals.mpw0, unweighted
    // acceptance, products through Species::addParticle, neutrals never depleted.  A ch4/v2 main selects that algorithm here.
    void useFixedWeights(bool on) { if (picg_mcc_set_variant(handle.get(), on ? 1 : 0) != PICG_OK) throw std::runtime_error(picg_last_error()); }
};
#endif
