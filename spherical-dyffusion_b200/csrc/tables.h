// Host-side fp64 precompute: quadrature, orthonormal associated Legendre tables, longitude DFT bases.
// Replaces torch_harmonics' quadrature.py / legendre.py precompute used at sfnonet.py:551-554
// (algorithm restated from SURVEY.md Appendix A; torch-harmonics itself is not in the reference tree).
#pragma once
#include <vector>

namespace sfno {

// nodes ascending in cos(theta) on [-1,1]
void legendre_gauss(int n, std::vector<double>& nodes, std::vector<double>& weights);
void clenshaw_curtis(int n, std::vector<double>& nodes, std::vector<double>& weights);

struct ShtTables {
  int nlat = 0, nlon = 0, lmax = 0, mmax = 0, grid = 0;
  std::vector<double> cost;     // [nlat] cos(colatitude), north -> south
  std::vector<double> quad_w;   // [nlat]
  std::vector<double> pct;      // [mmax][lmax][nlat] synthesis table
  std::vector<double> weights;  // [mmax][lmax][nlat] analysis table (pct * quad_w)
};
// returns false on invalid arguments
bool build_sht_tables(int nlat, int nlon, int lmax, int mmax, int grid, ShtTables& out);

// forward DFT basis  E[n = 2m+ri][j] = (2 pi / nlon) * {cos, -sin}(2 pi m j / nlon),  n < 2*mmax
void build_dft_forward(int nlon, int mmax, std::vector<double>& e);
// inverse (C2R, unnormalised) basis  E[j][kk = 2m+ri] = c_m * {cos, -sin}(2 pi m j / nlon)
void build_dft_inverse(int nlon, int mmax, std::vector<double>& e);

}  // namespace sfno
