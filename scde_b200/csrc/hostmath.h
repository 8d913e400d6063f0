// hostmath.h -- host-side helpers (see hostmath.cpp)
#pragma once
#include <cstdint>

namespace scde {

class GlibcRand {
public:
    explicit GlibcRand(uint32_t seed);
    int32_t next();        // == rand()
    int draw(int n);       // == while (n <= (rj = rand() / (RAND_MAX / n))); rj
private:
    int32_t r_[31];
    int f_, b_;
};

void bh_cz(const double *z, int n, double *cz);

}  // namespace scde
