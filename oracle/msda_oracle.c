/*
 * TEST INFRASTRUCTURE ONLY.  CPU oracle for multi-scale deformable attention.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker.  The product path
 * (snipper_b200/) never links or imports it.
 *
 * Parity is PINNED: oracle/make_golden.py runs the reference's own
 * ms_deform_attn_core_pytorch (models/ops/functions/ms_deform_attn_func.py:45-65) and
 * MSDeformAttn module (models/ops/modules/ms_deform_attn.py:99-243) in the build
 * container and commits their outputs under tests/golden/; tests/test_oracle_golden.py
 * checks this file against those vectors.
 *
 * Build: make -C oracle   (gcc -O2 -shared -fPIC, see oracle/Makefile)
 */
#include <math.h>
#include <stdint.h>

#define REAL double
#define SUFFIX f64
#define FLOOR floor
#define FMA fma
#include "msda_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR
#undef FMA

#define REAL float
#define SUFFIX f32
#define FLOOR floorf
#define FMA fmaf
#include "msda_oracle_body.inc"
#undef REAL
#undef SUFFIX
#undef FLOOR
#undef FMA

int msda_oracle_abi_version(void) { return 1; }
