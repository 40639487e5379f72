// refshim forwarding header (TEST INFRASTRUCTURE ONLY, see refshim_pcl.h)
#pragma once
#include "../../../refshim_pcl.h"
