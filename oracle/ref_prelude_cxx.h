// TEST INFRASTRUCTURE ONLY.
// Force-included (-include) in front of the UNMODIFIED reference translation units
// (/root/reference/spmv-csrk/csrk.cpp and our shim) when building oracle/_ref.
// It pulls in every system header the reference uses *before* any macro games, so that
// the macros below only rewrite the reference's own text:
//   * REF_REAL_DOUBLE : `float` -> `double`.  The reference kernels are fp32
//     (spmv-csrk/csrk.h:40 `typedef float val_type`); north_star asks for fp64, so the
//     fp64 oracle is the same source with the scalar type widened.
//   * __clang__ : selects the k=2 loop at spmv-csrk/csrk.cpp:167-223; the non-clang branch
//     has a debug printf per super-row (csrk.cpp:231).
#pragma once
#include <algorithm>
#include <cassert>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <map>
#include <omp.h>
#include <set>
#include <string>
#include <sys/time.h>
#include <vector>
#include <boost/graph/adjacency_list.hpp>
#include <boost/graph/sequential_vertex_coloring.hpp>
#ifdef REF_REAL_DOUBLE
#define float double
#endif
#ifndef __clang__
#define __clang__ 1
#endif
