// TEST INFRASTRUCTURE ONLY -- forwards the engine include path "editor/editor_node.h" to the stand-in (see godot_shim.h)
#pragma once
#include "../godot_shim.h"
