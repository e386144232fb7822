// forwards to the host-side io mirror (see ../mrmd_b200_io.hpp)
#pragma once
#include "../mrmd_b200_io.hpp"
