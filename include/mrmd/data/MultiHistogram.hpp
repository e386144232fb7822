// data/MultiHistogram.hpp of the reference: the class lives in mrmd_b200.hpp
#pragma once
#include "../mrmd_b200.hpp"
