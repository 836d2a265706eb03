/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.
 *
 * CPU oracle for the TUCH self-contact hot path: a scalar restatement of the reference's
 * PyTorch tensor algebra (tuch/utils/contact.py, the masked nearest-vertex search of
 * tuch/smplify/losses.py:76-93 and the region minimum of losses.py:113-116).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library; the product (tuch_b200/) never does.
 *
 * Parity pin: validated against outputs of the reference's own Python functions recorded in
 * tests/golden/ by tests/golden/make_golden.py (tests/test_oracle_golden.py).
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stddef.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL float
#define FN(name) CAT(name, _f32)
#define SQRT sqrtf
#define ATAN2 atan2f
#include "contact_oracle_impl.h"
#undef REAL
#undef FN
#undef SQRT
#undef ATAN2

#define REAL double
#define FN(name) CAT(name, _f64)
#define SQRT sqrt
#define ATAN2 atan2
#include "contact_oracle_impl.h"
#undef REAL
#undef FN
#undef SQRT
#undef ATAN2

int oracle_num_threads(void)
{
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
