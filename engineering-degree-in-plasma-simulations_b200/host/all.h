// Facade header mirroring ch4/v3/src/all.h: scalar type, debug macro, physical constants.
// The constants are the reference's literal values (all.h:13-23; note the 10-digit pi) because they
// enter the parity-critical arithmetic.
#ifndef ALL_H
#define ALL_H
#include <iostream>

#ifdef DEBUG
#define dmsg(x) std::cerr << x
#else
#define dmsg(x)
#endif

using type_calc = double;   // the device path computes in fp64 (all.h:11)

namespace Const {
const double eps_0 = 8.85418782e-12;    // C/(V*m)
const double q_e = 1.602176565e-19;     // C
const double amu = 1.660538921e-27;     // kg
const double m_e = 9.10938215e-31;      // kg
const double k = 1.380648e-23;          // J/K
const double pi = 3.141592653;
const double eV_to_K = q_e / k;
const double N_a = 6.02214076e23;
}  // namespace Const
#endif
