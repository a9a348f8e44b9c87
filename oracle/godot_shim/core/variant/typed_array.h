// TEST INFRASTRUCTURE ONLY -- forwards the engine include path "core/variant/typed_array.h" to the stand-in (see godot_shim.h)
#pragma once
#include "../../godot_shim.h"
