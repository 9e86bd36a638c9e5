// Facade of ch4/v3/src/Interactions.h: the Interaction interface and MC_MEX_Ionization, the one interaction the
// v3 main loop instantiates (main.cpp:122).  apply(dt) runs the per-cell Monte-Carlo kernel (csrc/mcc.cu).
#ifndef INTERACTIONS_H
#define INTERACTIONS_H
#include <memory>
#include <string>
#include "Species.h"
#include "World.h"
#include "picgpu.h"

class Interaction {
public:
    virtual void apply(type_calc dt) noexcept = 0;
    virtual ~Interaction() noexcept = default;
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
};
#endif
