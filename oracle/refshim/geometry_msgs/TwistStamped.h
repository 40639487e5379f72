// refshim forwarding header (TEST INFRASTRUCTURE ONLY, see refshim_ros.h)
#pragma once
#include "../refshim_ros.h"
