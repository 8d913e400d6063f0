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

// The tail of R's density.default after BinDist: y[n] = binned mass on the working grid seq(lo, up, length = n) (the upper
// half of the 2n-point FFT buffer is zero) -> circular correlation with the Gaussian kernel of s.d. bw on 2n points,
// pmax(0, .), linear interpolation (approx) onto seq(from, to, length = n_user).
void density_from_bins(const double *y, int n, double lo, double up, double bw, int n_user, double from, double to,
                       double *xout, double *yout);

}  // namespace scde
