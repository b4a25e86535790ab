/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference algorithm.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function here
 * against (a) the committed golden vectors in tests/golden/ that were
 * generated from the unmodified reference (tests/golden/make_golden.py) and
 * (b) the reference's own compiled core (oracle/_ref) wherever that .so is
 * present, byte for byte.
 *
 * Nothing in the product package (adrt_b200/) links, imports or executes
 * this file.  Only tests/, bench.py's cpu_baseline/reference legs and
 * __graft_entry__.smoke() may.
 */
#include <math.h>
#include <stddef.h>

#define T float
#define FN(name) name##_f32
#include "adrt_oracle_impl.h"
#undef T
#undef FN

#define T double
#define FN(name) name##_f64
#include "adrt_oracle_impl.h"
#undef T
#undef FN
