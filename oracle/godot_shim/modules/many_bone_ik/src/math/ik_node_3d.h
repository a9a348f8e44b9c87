// TEST INFRASTRUCTURE ONLY -- the reference tests include their module as "modules/many_bone_ik/src/math/ik_node_3d.h";
// forwards to the same file found through -I/root/reference/src (nothing is copied)
#pragma once
#include <math/ik_node_3d.h>
