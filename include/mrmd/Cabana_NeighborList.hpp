// the reference's drivers include this Cabana header for HalfNeighborList / FullNeighborList: see mrmd_b200.hpp
#pragma once
#include "mrmd_b200.hpp"
