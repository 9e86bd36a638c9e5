// Facade of ch4/v3/src/Outputs.h (I/O is outside the hot path): fieldsOutput writes the reference's .vti arrays as appended raw
// binary straight from the device, the CSV writers are kept minimal so that the main loop runs unchanged.
#ifndef _OUTPUT_H
#define _OUTPUT_H
#include <fstream>
#include <string>
#include <vector>
#include "Species.h"
#include "World.h"
#include "funkc.h"

namespace Output {
enum modes { none, all, screen, fields, particles, diagnostics, convergence };
void fieldsOutput(World& world, std::vector<Species>& species, std::string name1 = "");
void screenOutput(World& world, std::vector<Species>& species);
void diagOutput(World& world, std::vector<Species>& species);
void particlesOutput(World& world, std::vector<Species>& species, int num_parts_to_output_base, std::string name1 = "");
// restart files (not in the reference): world fields + particle stores + averages + RNG stream positions, see picgpu.h.
// load expects World and species rebuilt as at start-up; both return false (and log to cerr) on failure.
bool saveCheckpoint(const std::string& path, World& world, std::vector<Species>& species);
bool loadCheckpoint(const std::string& path, World& world, std::vector<Species>& species);
std::ostream& operator<<(std::ostream& out, Output::modes& type);
std::istream& operator>>(std::istream& in, Output::modes& type);
}  // namespace Output
#endif
