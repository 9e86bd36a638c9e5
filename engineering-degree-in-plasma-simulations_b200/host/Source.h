// Facade of ch4/v3/src/Source.h: face inlet sources; sample() injects on the device (csrc/source.cu).
#ifndef SOURCE_H
#define SOURCE_H
#include <memory>
#include <string>
#include "Species.h"
#include "World.h"
#include "picgpu.h"

class Source {
protected:
    Species& sp;
    World& world;
    std::shared_ptr<picg_source_s> handle;
    Source(Species& species, World& world, type_calc v_drift, type_calc den, type_calc T, std::string inlet_face) noexcept;

public:
    virtual ~Source() noexcept = default;
    virtual void sample() const noexcept;
};
class ColdBeamSource : public Source {
public:
    ColdBeamSource(Species& species, World& world, type_calc v_drift, type_calc den, std::string inlet_face = "-z", type_calc area_frac = 1.0) noexcept
        : Source(species, world, v_drift, den, 0.0, inlet_face) { (void)area_frac; }
};
class WarmBeamSource : public Source {
public:
    WarmBeamSource(Species& species, World& world, type_calc v_drift, type_calc den, type_calc T, std::string inlet_face = "-z", type_calc area_frac = 1.0) noexcept
        : Source(species, world, v_drift, den, T, inlet_face) { (void)area_frac; }
};
#endif
