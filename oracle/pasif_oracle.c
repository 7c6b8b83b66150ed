/* pasif_oracle.c -- plain-C CPU oracle of the reference's RK4 forward / adjoint hot path.
 * TEST INFRASTRUCTURE ONLY: loaded by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs as the checker / reported CPU baseline; never by the product path (pasif_b200/).
 * Validated against the NumPy restatement (oracle/pasif_oracle.py) and against golden vectors produced by the
 * compiled, unmodified reference on a B200 (tests/golden/ref_*.npz). */
#include <stdlib.h>
#include <string.h>

#define REAL double
#define NAME(x) x##_f64
#include "pasif_oracle_impl.h"
#undef REAL
#undef NAME

#define REAL float
#define NAME(x) x##_f32
#include "pasif_oracle_impl.h"
#undef REAL
#undef NAME
