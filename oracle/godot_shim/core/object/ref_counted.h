// TEST INFRASTRUCTURE ONLY -- forwards the engine include path "core/object/ref_counted.h" to the stand-in (see godot_shim.h)
#pragma once
#include "../../godot_shim.h"
