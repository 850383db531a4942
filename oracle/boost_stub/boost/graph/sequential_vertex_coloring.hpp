// TEST INFRASTRUCTURE ONLY -- see adjacency_list.hpp in this directory.
#pragma once
#include "adjacency_list.hpp"
