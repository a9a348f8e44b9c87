// TEST INFRASTRUCTURE ONLY -- forwards the engine include path "scene/3d/node_3d.h" to the stand-in (see godot_shim.h)
#pragma once
#include "../../godot_shim.h"
