// forwards to the single-file mirror of mrmd::io (see ../mrmd_b200_io.hpp)
#pragma once
#include "../mrmd_b200_io.hpp"
