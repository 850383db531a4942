/* TEST INFRASTRUCTURE ONLY.
 * Force-included in front of the UNMODIFIED /root/reference/spmv-csr/spmv.c when building
 * oracle/_ref/libref_csr_{f32,f64}.so.  System headers first, then:
 *   - rename the driver's main() so the file can live in a shared library;
 *   - REF_REAL_DOUBLE widens `float` to `double` (the reference's spmv.h:8-12 declares the
 *     double signatures, its spmv.c:12 implements float).  With the widening, the file's own
 *     my_read_csr ("%f" into a double) must NOT be used -- the fp64 library is only entered
 *     through omp_spmv / test_spmv. */
#pragma once
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#define main ref_spmv_csr_main
#ifdef REF_REAL_DOUBLE
#define float double
#endif
