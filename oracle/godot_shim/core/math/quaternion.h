// TEST INFRASTRUCTURE ONLY -- forwards the engine include path "core/math/quaternion.h" to the stand-in (see godot_shim.h)
#pragma once
#include "../../godot_shim.h"
