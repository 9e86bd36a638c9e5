// Facade of ch4/v3/src/Outputs.h (I/O is outside the hot path; kept minimal so that the main loop runs unchanged).
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
std::ostream& operator<<(std::ostream& out, Output::modes& type);
std::istream& operator>>(std::istream& in, Output::modes& type);
}  // namespace Output
#endif
