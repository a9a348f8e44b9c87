// TEST INFRASTRUCTURE ONLY -- the reference tests include their module as "modules/many_bone_ik/src/math/qcp.h";
// forwards to the same file found through -I/root/reference/src (nothing is copied)
#pragma once
#include <math/qcp.h>
