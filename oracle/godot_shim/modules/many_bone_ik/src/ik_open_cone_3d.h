// TEST INFRASTRUCTURE ONLY -- the reference tests include their module as "modules/many_bone_ik/src/ik_open_cone_3d.h";
// forwards to the same file found through -I/root/reference/src (nothing is copied)
#pragma once
#include <ik_open_cone_3d.h>
