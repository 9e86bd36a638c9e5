// Facade of ch4/v3/src/Rnd.h.  `rnd = Rnd(seed);` (the way a harness pins the reference, Rnd.cpp:7) also seeds the
// Philox streams of every stochastic device kernel (picg_seed).  Host draws keep std::mt19937.
#ifndef RND_H
#define RND_H
#include <random>
#include "all.h"

class Rnd {
protected:
    std::mt19937 mt_gen;
    std::uniform_real_distribution<type_calc> rnd_dist;

public:
    Rnd();
    Rnd(unsigned seed);
    type_calc operator()() { return rnd_dist(mt_gen); }
    type_calc operator()(type_calc min, type_calc max) { return min + rnd_dist(mt_gen) * (max - min); }
};
extern Rnd rnd;
#endif
