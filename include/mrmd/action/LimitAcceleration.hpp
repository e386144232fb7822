// forwards to the single-file mirror of the reference API (see ../mrmd_b200.hpp)
#pragma once
#include "../mrmd_b200.hpp"
