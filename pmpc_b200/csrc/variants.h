// The compiled kernel table: X(NX, NU, KIND, VAR) — see inst.cu.  Keep in sync with VARIANTS in the Makefile.
#pragma once
#define PMPC_VARIANTS(X)                                                                              \
  X(2, 1, 0, 0) X(2, 1, 0, 2) X(2, 1, 1, 0) X(2, 1, 1, 2) X(2, 1, 2, 0) X(2, 1, 2, 2)                 \
  X(4, 2, 0, 0) X(4, 2, 0, 1) X(4, 2, 0, 2) X(4, 2, 0, 3) X(4, 2, 1, 0) X(4, 2, 1, 1) X(4, 2, 1, 2)   \
  X(4, 2, 1, 3) X(4, 2, 2, 0) X(4, 2, 2, 1) X(4, 2, 2, 2) X(4, 2, 2, 3)                               \
  X(8, 4, 0, 0) X(8, 4, 1, 0) X(8, 4, 2, 0) X(12, 4, 0, 0) X(12, 4, 1, 0) X(12, 4, 2, 0)              \
  X(16, 8, 0, 0) X(16, 8, 1, 0) X(16, 8, 2, 0) X(32, 8, 0, 0) X(32, 8, 1, 0) X(32, 8, 2, 0)
