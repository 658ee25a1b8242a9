// Device-side SHT plan: Legendre and DFT tables in the layouts the GEMM ops consume.
#pragma once
#include <vector>

#include "common.cuh"
#include "tables.h"

namespace sfno {

struct ShtDeviceTables {
  int nlat = 0, nlon = 0, lmax = 0, mmax = 0, grid = 0, precision = 0;
  int Kp = 0;   // nlat rounded up to 8  (row length of F/G rows and of wq rows)
  int Lq = 0;   // lmax rounded up to 8  (row length of pt rows)
  int Wp = 0;   // nlon rounded up to 8  (row length of efwd rows)
  int Kq2 = 0;  // 2*mmax rounded up to 8 (row length of einv rows)
  void* wq = nullptr;    // [mmax][lmax][Kp]   analysis  (Legendre x quadrature weight), A operand of OpLeg
  void* pt = nullptr;    // [mmax][nlat][Lq]   synthesis, transposed: B operand of OpIleg
  int basis_reps = 16;   // replicas of the two DFT bases (L2 broadcast spreading)
  void* efwd = nullptr;  // [basis_reps][2*mmax][Wp]  forward DFT basis, A operand of OpDft
  void* einv = nullptr;  // [basis_reps][nlon][Kq2]   inverse DFT basis, B operand of OpIdft
  // transposed tables of the ADJOINT transforms (backward pass, built on demand by sht_tables_enable_adjoint):
  //   adjoint of the analysis  = synthesis-shaped ops with (wq^T, efwd^T);  adjoint of the synthesis = analysis-shaped ops
  //   with (einv^T, pct in [m][l][k] layout)
  void* wq_t = nullptr;    // [mmax][nlat][Lq]            analysis table transposed   (B operand of OpIleg)
  void* efwd_t = nullptr;  // [basis_reps][nlon][Kq2]     forward DFT basis transposed (B operand of OpIdft)
  void* pct_a = nullptr;   // [mmax][lmax][Kp]            synthesis table in analysis layout (A operand of OpLeg)
  void* einv_t = nullptr;  // [basis_reps][2*mmax][Wp]    inverse DFT basis transposed (A operand of OpDft)
  size_t bytes = 0;
};

int sht_tables_upload(int nlat, int nlon, int lmax, int mmax, int grid, int precision, ShtDeviceTables& out);
int sht_tables_enable_adjoint(ShtDeviceTables& t);
void sht_tables_free(ShtDeviceTables& t);

}  // namespace sfno

struct sfno_sht_plan {
  sfno::ShtDeviceTables t;
};
