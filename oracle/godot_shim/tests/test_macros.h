// TEST INFRASTRUCTURE ONLY -- stand-in for the engine's "tests/test_macros.h" (doctest): just enough to
// run the reference module's own test headers (/root/reference/tests/*.h, included unmodified by
// oracle/ref_doctests_main.cpp) against the reference sources compiled in oracle/_ref.
#pragma once
#include "../godot_shim.h"

#include <cstdio>
#include <vector>
// the reference's test_qcp.h calls an unqualified abs() on floats: the C++ <math.h>/<stdlib.h> wrappers put the
// floating-point overloads into the global namespace (with <cmath> alone only ::abs(int) is visible and every
// difference truncates to 0)
#include <math.h>
#include <stdlib.h>

namespace shim_doctest {
struct Case {
	const char *name;
	void (*fn)();
};
inline std::vector<Case> &cases() {
	static std::vector<Case> c;
	return c;
}
struct Counters {
	int checks = 0, failed = 0;
};
inline Counters &counters() {
	static Counters c;
	return c;
}
struct Registrar {
	Registrar(const char *p_name, void (*p_fn)()) { cases().push_back(Case{ p_name, p_fn }); }
};
struct RequireFailed {};
inline void report(bool ok, const char *kind, const char *expr, const char *file, int line) {
	counters().checks++;
	if (!ok) {
		counters().failed++;
		fprintf(stderr, "  FAILED %s(%s) at %s:%d\n", kind, expr, file, line);
	}
}
} // namespace shim_doctest

#define SHIM_DT_CAT2(a, b) a##b
#define SHIM_DT_CAT(a, b) SHIM_DT_CAT2(a, b)
#define TEST_CASE(m_name)                                                                                               \
	static void SHIM_DT_CAT(shim_case_, __LINE__)();                                                                    \
	static shim_doctest::Registrar SHIM_DT_CAT(shim_reg_, __LINE__)(m_name, &SHIM_DT_CAT(shim_case_, __LINE__));        \
	static void SHIM_DT_CAT(shim_case_, __LINE__)()
#define CHECK(...) shim_doctest::report(static_cast<bool>(__VA_ARGS__), "CHECK", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_FALSE(...) shim_doctest::report(!static_cast<bool>(__VA_ARGS__), "CHECK_FALSE", #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_MESSAGE(m_cond, ...) shim_doctest::report(static_cast<bool>(m_cond), "CHECK", #m_cond, __FILE__, __LINE__)
#define CHECK_EQ(m_a, m_b) shim_doctest::report((m_a) == (m_b), "CHECK_EQ", #m_a ", " #m_b, __FILE__, __LINE__)
#define CHECK_NE(m_a, m_b) shim_doctest::report((m_a) != (m_b), "CHECK_NE", #m_a ", " #m_b, __FILE__, __LINE__)
#define CHECK_LT(m_a, m_b) shim_doctest::report((m_a) < (m_b), "CHECK_LT", #m_a ", " #m_b, __FILE__, __LINE__)
#define CHECK_GT(m_a, m_b) shim_doctest::report((m_a) > (m_b), "CHECK_GT", #m_a ", " #m_b, __FILE__, __LINE__)
#define REQUIRE(...)                                                                                   \
	do {                                                                                               \
		bool shim_ok = static_cast<bool>(__VA_ARGS__);                                                 \
		shim_doctest::report(shim_ok, "REQUIRE", #__VA_ARGS__, __FILE__, __LINE__);                    \
		if (!shim_ok) {                                                                                \
			throw shim_doctest::RequireFailed();                                                       \
		}                                                                                              \
	} while (0)
