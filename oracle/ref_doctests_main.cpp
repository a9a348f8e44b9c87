// TEST INFRASTRUCTURE ONLY -- runs the reference module's OWN unit tests (/root/reference/tests/test_qcp.h,
// test_ik_node_3d.h, test_ik_kusudama_3d.h: 15 doctest cases, included unmodified from where they lie)
// against the reference's own sources compiled in oracle/_ref over the engine stand-in
// (oracle/godot_shim/, engine math = oracle/godot_math.h).  What this pins: the engine-math restatement
// and the stand-in object model reproduce every result the reference's tests assert.
// Built by `make -C oracle ref` only when /root/reference is present; prints one line per case and
// "ref_doctests: <cases> cases, <checks> checks, <failed> failed"; exit code = number of failed checks.
#include "test_ik_kusudama_3d.h"
#include "test_ik_node_3d.h"
#include "test_qcp.h"

int main() {
	int aborted = 0;
	for (const shim_doctest::Case &c : shim_doctest::cases()) {
		int before = shim_doctest::counters().failed;
		try {
			c.fn();
		} catch (const shim_doctest::RequireFailed &) {
			aborted++;
		}
		printf("%s %s\n", shim_doctest::counters().failed == before ? "ok    " : "FAILED", c.name);
	}
	printf("ref_doctests: %d cases, %d checks, %d failed\n", (int)shim_doctest::cases().size(), shim_doctest::counters().checks,
			shim_doctest::counters().failed);
	return shim_doctest::counters().failed + aborted;
}
