// TEST INFRASTRUCTURE ONLY -- stand-in for the Godot engine headers the reference module includes.
//
// Purpose: compile the reference's OWN translation units (src/*.cpp, src/math/*.cpp, read where they
// lie under /root/reference, unmodified) into oracle/_ref/libmbik_ref.so, so that the restatement in
// oracle/ewbik_oracle.cpp and the CUDA path can be compared with the reference's real solver code.
// The engine itself (Godot 4.3/4.4) is not in the tree, so this file supplies the API surface the
// module touches:
//   * core/math   -> oracle/godot_math.h (the same restatement of Vector3/Basis/Quaternion/Transform3D
//                    the oracle uses; the engine arithmetic therefore stays "restated", the module's
//                    logic becomes "the real thing");
//   * containers  -> Vector<T> (copy-on-write like the engine's), List<T>, HashMap / RBSet over std::map;
//                    String / StringName / NodePath over std::string;
//   * object model-> Object/RefCounted/Ref<T>/WeakRef/Variant/ClassDB: reference counting and
//                    dynamic casts behave like the engine's; method binding, signals, property
//                    lists and editor hooks are inert;
//   * scene       -> Node/Node3D/Skeleton3D/SkeletonModifier3D: plain data holders with the engine's
//                    accessor names (bone poses, parents, children in ascending bone index, global
//                    transform of a node = its stored transform).
// Nothing here is copied from the engine or the reference; it is the minimum that makes the
// reference's code run headless.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <initializer_list>
#include <map>
#include <string>
#include <type_traits>
#include <vector>

#include "../godot_math.h"

using namespace gd;
using gd::real_t;

// ---------------------------------------------------------------- typedefs / macros
#ifndef MIN
#define MIN(m_a, m_b) (((m_a) < (m_b)) ? (m_a) : (m_b))
#endif
#ifndef MAX
#define MAX(m_a, m_b) (((m_a) > (m_b)) ? (m_a) : (m_b))
#endif
#ifndef CLAMP
#define CLAMP(m_a, m_min, m_max) (((m_a) < (m_min)) ? (m_min) : (((m_a) > (m_max)) ? m_max : m_a))
#endif
#ifndef ABS
#define ABS(m_v) (((m_v) < 0) ? (-(m_v)) : (m_v))
#endif
#ifndef SIGN
#define SIGN(m_v) (((m_v) > 0) ? (1.0f) : (((m_v) < 0) ? (-1.0f) : (0.0f)))
#endif
#define _FORCE_INLINE_ inline
#define _ALWAYS_INLINE_ inline
#define likely(x) (x)
#define unlikely(x) (x)

namespace gd {
namespace Math {
inline float tan(float x) { return ::tanf(x); }
inline double tan(double x) { return ::tan(x); }
inline float atan2(float y, float x) { return ::atan2f(y, x); }
inline double atan2(double y, double x) { return ::atan2(y, x); }
inline float asin(float x) { return x < -1.0f ? (float)(-Math_PI / 2) : (x > 1.0f ? (float)(Math_PI / 2) : ::asinf(x)); }
inline double asin(double x) { return x < -1.0 ? (-Math_PI / 2) : (x > 1.0 ? (Math_PI / 2) : ::asin(x)); }
inline float pow(float x, float y) { return ::powf(x, y); }
inline double pow(double x, double y) { return ::pow(x, y); }
inline float floor(float x) { return ::floorf(x); }
inline double floor(double x) { return ::floor(x); }
inline float fmod(float x, float y) { return ::fmodf(x, y); }
inline double fmod(double x, double y) { return ::fmod(x, y); }
inline float rad_to_deg(float r) { return r * (float)(180.0 / Math_PI); }
inline double rad_to_deg(double r) { return r * (180.0 / Math_PI); }
inline double deg_to_rad(double d) { return d * (Math_PI / 180.0); }
inline bool is_inf(float x) { return std::isinf(x); }
inline bool is_inf(double x) { return std::isinf(x); }
inline bool is_finite(double x) { return std::isfinite(x); }
inline double lerp(double a, double b, double w) { return a + (b - a) * w; }
inline bool is_equal_approx(double a, double b) {
	if (a == b) {
		return true;
	}
	double tolerance = CMP_EPSILON * abs(a);
	if (tolerance < CMP_EPSILON) {
		tolerance = CMP_EPSILON;
	}
	return abs(a - b) < tolerance;
}
inline float fposmod(float x, float y) {
	float v = ::fmodf(x, y);
	if (((v < 0) && (y > 0)) || ((v > 0) && (y < 0))) {
		v += y;
	}
	v += 0.0f;
	return v;
}
} // namespace Math
} // namespace gd

#define Math_SQRT12 0.7071067811865475244008443621048490
#define Math_SQRT2 1.4142135623730950488016887242

// ---------------------------------------------------------------- error macros (log-less: the engine logs and returns)
#define ERR_FAIL_COND(m_cond) \
	if (m_cond) {             \
		return;               \
	} else                    \
		((void)0)
#define ERR_FAIL_COND_MSG(m_cond, m_msg) ERR_FAIL_COND(m_cond)
#define ERR_FAIL_COND_V(m_cond, m_ret) \
	if (m_cond) {                      \
		return m_ret;                  \
	} else                             \
		((void)0)
#define ERR_FAIL_COND_V_MSG(m_cond, m_ret, m_msg) ERR_FAIL_COND_V(m_cond, m_ret)
#define ERR_FAIL_NULL(m_p) ERR_FAIL_COND((m_p) == nullptr)
#define ERR_FAIL_NULL_MSG(m_p, m_msg) ERR_FAIL_COND((m_p) == nullptr)
#define ERR_FAIL_NULL_V(m_p, m_ret) ERR_FAIL_COND_V((m_p) == nullptr, m_ret)
#define ERR_FAIL_NULL_V_MSG(m_p, m_ret, m_msg) ERR_FAIL_COND_V((m_p) == nullptr, m_ret)
#define ERR_FAIL_INDEX(m_i, m_n) ERR_FAIL_COND((int64_t)(m_i) < 0 || (int64_t)(m_i) >= (int64_t)(m_n))
#define ERR_FAIL_INDEX_MSG(m_i, m_n, m_msg) ERR_FAIL_INDEX(m_i, m_n)
#define ERR_FAIL_INDEX_V(m_i, m_n, m_ret) ERR_FAIL_COND_V((int64_t)(m_i) < 0 || (int64_t)(m_i) >= (int64_t)(m_n), m_ret)
#define ERR_FAIL_INDEX_V_MSG(m_i, m_n, m_ret, m_msg) ERR_FAIL_INDEX_V(m_i, m_n, m_ret)
#define ERR_FAIL() return
#define ERR_FAIL_MSG(m_msg) return
#define ERR_FAIL_V(m_ret) return m_ret
#define ERR_FAIL_V_MSG(m_ret, m_msg) return m_ret
#define ERR_CONTINUE(m_cond) \
	if (m_cond) {            \
		continue;            \
	} else                   \
		((void)0)
#define ERR_CONTINUE_MSG(m_cond, m_msg) ERR_CONTINUE(m_cond)
#define ERR_BREAK(m_cond) \
	if (m_cond) {         \
		break;            \
	} else                \
		((void)0)
#define ERR_PRINT(m_msg) ((void)0)
#define ERR_PRINT_ONCE(m_msg) ((void)0)
#define WARN_PRINT(m_msg) ((void)0)
#define WARN_PRINT_ONCE(m_msg) ((void)0)
#define DEV_ASSERT(m_cond) ((void)0)
#define CRASH_COND(m_cond) \
	if (m_cond) {          \
		abort();           \
	} else                 \
		((void)0)
#define CRASH_COND_MSG(m_cond, m_msg) CRASH_COND(m_cond)

// ---------------------------------------------------------------- small math types outside godot_math.h
struct Vector2 {
	real_t x = 0, y = 0;
	Vector2() {}
	Vector2(real_t p_x, real_t p_y) : x(p_x), y(p_y) {}
	bool operator==(const Vector2 &o) const { return x == o.x && y == o.y; }
	bool operator!=(const Vector2 &o) const { return !(*this == o); }
};
struct Vector4 {
	real_t x = 0, y = 0, z = 0, w = 0;
	Vector4() {}
	Vector4(real_t p_x, real_t p_y, real_t p_z, real_t p_w) : x(p_x), y(p_y), z(p_z), w(p_w) {}
	bool operator==(const Vector4 &o) const { return x == o.x && y == o.y && z == o.z && w == o.w; }
	bool operator!=(const Vector4 &o) const { return !(*this == o); }
};
struct Color {
	float r = 0, g = 0, b = 0, a = 1;
	Color() {}
	Color(float p_r, float p_g, float p_b, float p_a = 1.0f) : r(p_r), g(p_g), b(p_b), a(p_a) {}
};

// ---------------------------------------------------------------- strings
class String {
	std::string s;

public:
	String() {}
	String(const char *p) : s(p ? p : "") {}
	String(const std::string &p) : s(p) {}
	String(const wchar_t *p) {
		for (; p && *p; p++) {
			s.push_back((char)*p);
		}
	}
	String operator+(double v) const { return String(s + std::to_string(v)); }
	const std::string &std_str() const { return s; }
	bool is_empty() const { return s.empty(); }
	int length() const { return (int)s.size(); }
	bool operator==(const String &o) const { return s == o.s; }
	bool operator!=(const String &o) const { return s != o.s; }
	bool operator==(const char *o) const { return s == o; }
	bool operator!=(const char *o) const { return s != o; }
	bool operator<(const String &o) const { return s < o.s; }
	String operator+(const String &o) const { return String(s + o.s); }
	String &operator+=(const String &o) {
		s += o.s;
		return *this;
	}
	bool begins_with(const String &p) const { return s.compare(0, p.s.size(), p.s) == 0; }
	bool ends_with(const String &p) const { return s.size() >= p.s.size() && s.compare(s.size() - p.s.size(), p.s.size(), p.s) == 0; }
	int find(const String &p, int from = 0) const {
		size_t r = s.find(p.s, (size_t)from);
		return r == std::string::npos ? -1 : (int)r;
	}
	bool contains(const String &p) const { return s.find(p.s) != std::string::npos; }
	String substr(int from, int chars = -1) const {
		if (from < 0 || from >= (int)s.size()) {
			return String();
		}
		return String(s.substr((size_t)from, chars < 0 ? std::string::npos : (size_t)chars));
	}
	int get_slice_count(const String &p_splitter) const {
		if (s.empty() || p_splitter.s.empty()) {
			return 0;
		}
		int n = 1;
		size_t pos = 0;
		while ((pos = s.find(p_splitter.s, pos)) != std::string::npos) {
			n++;
			pos += p_splitter.s.size();
		}
		return n;
	}
	String get_slice(const String &p_splitter, int p_slice) const {
		if (s.empty() || p_splitter.s.empty() || p_slice < 0) {
			return String();
		}
		size_t pos = 0;
		int i = 0;
		while (true) {
			size_t nx = s.find(p_splitter.s, pos);
			if (i == p_slice) {
				return String(s.substr(pos, nx == std::string::npos ? std::string::npos : nx - pos));
			}
			if (nx == std::string::npos) {
				return String();
			}
			pos = nx + p_splitter.s.size();
			i++;
		}
	}
	String get_slicec(char p_splitter, int p_slice) const { return get_slice(String(std::string(1, p_splitter)), p_slice); }
	int64_t to_int() const { return s.empty() ? 0 : strtoll(s.c_str(), nullptr, 10); }
	double to_float() const { return s.empty() ? 0.0 : strtod(s.c_str(), nullptr); }
	bool is_valid_int() const {
		if (s.empty()) {
			return false;
		}
		size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
		if (i >= s.size()) {
			return false;
		}
		for (; i < s.size(); i++) {
			if (s[i] < '0' || s[i] > '9') {
				return false;
			}
		}
		return true;
	}
	const char *utf8_ptr() const { return s.c_str(); }
};
inline String operator+(const char *a, const String &b) { return String(a) + b; }
inline String itos(int64_t v) { return String(std::to_string(v)); }
inline String rtos(double v) { return String(std::to_string(v)); }
template <typename... A>
inline String vformat(const String &fmt, A...) { return fmt; }
template <typename... A>
inline void print_line(A...) {}
template <typename... A>
inline void print_error(A...) {}
template <typename... A>
inline void print_verbose(A...) {}

class StringName : public String {
public:
	StringName() {}
	StringName(const char *p) : String(p) {}
	StringName(const String &p) : String(p) {}
	operator String() const { return String(std_str()); }
};
#define SNAME(m_s) StringName(m_s)
#define SceneStringName(m_s) StringName(#m_s)

class NodePath {
	String path;

public:
	NodePath() {}
	NodePath(const char *p) : path(p) {}
	NodePath(const String &p) : path(p) {}
	bool is_empty() const { return path.is_empty(); }
	operator String() const { return path; }
	bool operator==(const NodePath &o) const { return path == o.path; }
	bool operator!=(const NodePath &o) const { return path != o.path; }
	const String &str() const { return path; }
	StringName get_concatenated_names() const { return StringName(path); }
	int get_name_count() const { return path.is_empty() ? 0 : path.get_slice_count("/"); }
	StringName get_name(int i) const { return StringName(path.get_slice("/", i)); }
};

class StringBuilder {
	String acc;

public:
	StringBuilder &append(const String &p) {
		acc += p;
		return *this;
	}
	StringBuilder &operator+(const String &p) { return append(p); }
	String as_string() const { return acc; }
};

// ---------------------------------------------------------------- Vector<T>: the engine's copy-on-write vector
// Copies share the buffer (reference count); the first write through a shared handle detaches it.  Same
// observable semantics and the same cost profile as the engine's CowData-backed Vector.
#include <memory>
template <typename T>
class Vector {
	std::shared_ptr<std::vector<T>> d;
	std::vector<T> &w() {
		if (!d) {
			d = std::make_shared<std::vector<T>>();
		} else if (d.use_count() > 1) {
			d = std::make_shared<std::vector<T>>(*d);
		}
		return *d;
	}
	const std::vector<T> &r() const {
		static const std::vector<T> empty;
		return d ? *d : empty;
	}

public:
	struct Write {
		Vector *owner;
		T &operator[](int64_t i) { return owner->w()[(size_t)i]; }
	} write;

	Vector() { write.owner = this; }
	Vector(const Vector &o) : d(o.d) { write.owner = this; }
	Vector(std::initializer_list<T> l) : d(std::make_shared<std::vector<T>>(l)) { write.owner = this; }
	Vector &operator=(const Vector &o) {
		d = o.d;
		write.owner = this;
		return *this;
	}
	int64_t size() const { return (int64_t)r().size(); }
	bool is_empty() const { return r().empty(); }
	int resize(int64_t n) {
		w().resize((size_t)n);
		return 0;
	}
	void clear() { d.reset(); }
	bool push_back(const T &e) {
		T copy = e; // e may live in this buffer
		w().push_back(copy);
		return false;
	}
	void append(const T &e) { push_back(e); }
	void append_array(const Vector<T> &o) {
		std::vector<T> src = o.r();
		std::vector<T> &v = w();
		v.insert(v.end(), src.begin(), src.end());
	}
	void fill(const T &e) {
		T copy = e;
		std::vector<T> &v = w();
		std::fill(v.begin(), v.end(), copy);
	}
	void remove_at(int64_t i) {
		std::vector<T> &v = w();
		v.erase(v.begin() + i);
	}
	bool erase(const T &e) {
		int64_t i = find(e);
		if (i < 0) {
			return false;
		}
		remove_at(i);
		return true;
	}
	int insert(int64_t pos, const T &e) {
		T copy = e;
		std::vector<T> &v = w();
		v.insert(v.begin() + pos, copy);
		return 0;
	}
	void reverse() {
		std::vector<T> &v = w();
		std::reverse(v.begin(), v.end());
	}
	int64_t find(const T &e, int64_t from = 0) const {
		const std::vector<T> &v = r();
		for (int64_t i = from; i < (int64_t)v.size(); i++) {
			if (v[(size_t)i] == e) {
				return i;
			}
		}
		return -1;
	}
	bool has(const T &e) const { return find(e) >= 0; }
	void set(int64_t i, const T &e) {
		T copy = e;
		w()[(size_t)i] = copy;
	}
	const T &get(int64_t i) const { return r()[(size_t)i]; }
	const T &operator[](int64_t i) const { return r()[(size_t)i]; }
	const T *ptr() const { return r().data(); }
	T *ptrw() { return w().data(); }
	Vector<T> duplicate() const { return *this; }
	typedef typename std::vector<T>::iterator Iterator;
	typedef typename std::vector<T>::const_iterator ConstIterator;
	// a non-const range-for may write through the reference: detach first, as the engine's begin() does
	Iterator begin() { return w().begin(); }
	Iterator end() { return w().end(); }
	ConstIterator begin() const { return r().begin(); }
	ConstIterator end() const { return r().end(); }
	bool operator==(const Vector &o) const { return r() == o.r(); }
	bool operator!=(const Vector &o) const { return !(r() == o.r()); }
};
// std::vector<bool> has no data(); the module does not use Vector<bool>.
typedef Vector<Vector3> PackedVector3Array;
typedef Vector<int32_t> PackedInt32Array;
typedef Vector<float> PackedFloat32Array;
typedef Vector<double> PackedFloat64Array;
typedef Vector<String> PackedStringArray;
template <typename T>
using LocalVector = Vector<T>;

// ---------------------------------------------------------------- List<T>
template <typename T>
class List {
public:
	class Element {
		friend class List<T>;
		T value;
		Element *next_ptr = nullptr;
		Element *prev_ptr = nullptr;

	public:
		Element *next() { return next_ptr; }
		const Element *next() const { return next_ptr; }
		Element *prev() { return prev_ptr; }
		const Element *prev() const { return prev_ptr; }
		T &get() { return value; }
		const T &get() const { return value; }
	};

private:
	Element *first = nullptr;
	Element *last = nullptr;
	int count = 0;

public:
	List() {}
	List(const List &o) {
		for (const Element *e = o.first; e; e = e->next_ptr) {
			push_back(e->value);
		}
	}
	List &operator=(const List &o) {
		if (this != &o) {
			clear();
			for (const Element *e = o.first; e; e = e->next_ptr) {
				push_back(e->value);
			}
		}
		return *this;
	}
	~List() { clear(); }
	Element *front() { return first; }
	const Element *front() const { return first; }
	Element *back() { return last; }
	const Element *back() const { return last; }
	int size() const { return count; }
	bool is_empty() const { return count == 0; }
	Element *push_back(const T &v) {
		Element *e = new Element;
		e->value = v;
		e->prev_ptr = last;
		if (last) {
			last->next_ptr = e;
		} else {
			first = e;
		}
		last = e;
		count++;
		return e;
	}
	Element *push_front(const T &v) {
		Element *e = new Element;
		e->value = v;
		e->next_ptr = first;
		if (first) {
			first->prev_ptr = e;
		} else {
			last = e;
		}
		first = e;
		count++;
		return e;
	}
	Element *find(const T &v) {
		for (Element *e = first; e; e = e->next_ptr) {
			if (e->value == v) {
				return e;
			}
		}
		return nullptr;
	}
	bool erase(Element *e) {
		if (!e) {
			return false;
		}
		if (e->prev_ptr) {
			e->prev_ptr->next_ptr = e->next_ptr;
		} else {
			first = e->next_ptr;
		}
		if (e->next_ptr) {
			e->next_ptr->prev_ptr = e->prev_ptr;
		} else {
			last = e->prev_ptr;
		}
		delete e;
		count--;
		return true;
	}
	bool erase(const T &v) { return erase(find(v)); }
	void pop_front() { erase(first); }
	void pop_back() { erase(last); }
	void clear() {
		while (first) {
			erase(first);
		}
	}
	struct Iterator {
		Element *e;
		T &operator*() const { return e->get(); }
		T *operator->() const { return &e->get(); }
		Iterator &operator++() {
			e = e->next();
			return *this;
		}
		bool operator!=(const Iterator &o) const { return e != o.e; }
		bool operator==(const Iterator &o) const { return e == o.e; }
	};
	struct ConstIterator {
		const Element *e;
		const T &operator*() const { return e->get(); }
		const T *operator->() const { return &e->get(); }
		ConstIterator &operator++() {
			e = e->next();
			return *this;
		}
		bool operator!=(const ConstIterator &o) const { return e != o.e; }
		bool operator==(const ConstIterator &o) const { return e == o.e; }
	};
	Iterator begin() { return Iterator{ first }; }
	Iterator end() { return Iterator{ nullptr }; }
	ConstIterator begin() const { return ConstIterator{ first }; }
	ConstIterator end() const { return ConstIterator{ nullptr }; }
};

template <typename K, typename V>
class HashMap {
	std::map<K, V> m;

public:
	V &operator[](const K &k) { return m[k]; }
	const V &operator[](const K &k) const { return m.at(k); }
	bool has(const K &k) const { return m.find(k) != m.end(); }
	bool erase(const K &k) { return m.erase(k) > 0; }
	void clear() { m.clear(); }
	int size() const { return (int)m.size(); }
	void insert(const K &k, const V &v) { m[k] = v; }
	const V *getptr(const K &k) const {
		auto it = m.find(k);
		return it == m.end() ? nullptr : &it->second;
	}
};

template <typename T>
class RBSet {
	std::map<T, bool> m;

public:
	void insert(const T &k) { m[k] = true; }
	bool has(const T &k) const { return m.find(k) != m.end(); }
	bool erase(const T &k) { return m.erase(k) > 0; }
	int size() const { return (int)m.size(); }
};
template <typename T>
using HashSet = RBSet<T>;

// ---------------------------------------------------------------- object model
typedef int BoneId;
class ObjectID {
	uint64_t id = 0;

public:
	ObjectID() {}
	ObjectID(uint64_t p) : id(p) {}
	bool is_valid() const { return id != 0; }
	bool is_null() const { return id == 0; }
	operator uint64_t() const { return id; }
};
class Object;
class Variant;
class Array;
class Dictionary;
// A bound method with no arguments can be invoked (that is all the module's signal connections need:
// "modification_processed" -> ManyBoneIK3D::_update_ik_bones_transform, src/many_bone_ik_3d.cpp:1084).
class Callable {
public:
	std::function<void()> fn;
	const void *object = nullptr;
	std::string method_id; // raw bytes of the member pointer: identity for is_connected / disconnect
	Callable() {}
	template <typename... A>
	Callable bind(A...) const { return *this; }
	bool operator==(const Callable &o) const { return object == o.object && method_id == o.method_id; }
};
template <typename T, typename R, typename... P>
inline Callable callable_mp(T *p_instance, R (T::*p_method)(P...)) {
	Callable c;
	c.object = p_instance;
	c.method_id.assign(reinterpret_cast<const char *>(&p_method), sizeof(p_method));
	if constexpr (sizeof...(P) == 0) {
		c.fn = [p_instance, p_method]() { (p_instance->*p_method)(); };
	}
	return c;
}
template <typename T, typename R, typename... P>
inline Callable callable_mp(const T *p_instance, R (T::*p_method)(P...) const) {
	Callable c;
	c.object = p_instance;
	c.method_id.assign(reinterpret_cast<const char *>(&p_method), sizeof(p_method));
	if constexpr (sizeof...(P) == 0) {
		c.fn = [p_instance, p_method]() { (p_instance->*p_method)(); };
	}
	return c;
}
template <typename M>
inline Callable callable_mp_static(M) { return Callable(); }

struct PropertyInfo;
struct MethodInfo;

#define memnew(m_class) (new m_class)
template <typename T>
inline void memdelete(T *p) { delete p; }

enum PropertyHint {
	PROPERTY_HINT_NONE,
	PROPERTY_HINT_RANGE,
	PROPERTY_HINT_ENUM,
	PROPERTY_HINT_ENUM_SUGGESTION,
	PROPERTY_HINT_RESOURCE_TYPE,
	PROPERTY_HINT_NODE_PATH_VALID_TYPES,
	PROPERTY_HINT_NODE_TYPE,
	PROPERTY_HINT_ARRAY_TYPE,
	PROPERTY_HINT_TYPE_STRING,
	PROPERTY_HINT_FLAGS,
};
enum PropertyUsageFlags {
	PROPERTY_USAGE_NONE = 0,
	PROPERTY_USAGE_STORAGE = 1 << 1,
	PROPERTY_USAGE_EDITOR = 1 << 2,
	PROPERTY_USAGE_INTERNAL = 1 << 3,
	PROPERTY_USAGE_READ_ONLY = 1 << 28,
	PROPERTY_USAGE_ARRAY = 1 << 29,
	PROPERTY_USAGE_NO_EDITOR = PROPERTY_USAGE_STORAGE,
	PROPERTY_USAGE_DEFAULT = PROPERTY_USAGE_STORAGE | PROPERTY_USAGE_EDITOR,
};

class String;
class StringName;
class Object {
	std::map<std::string, std::vector<Callable>> signal_connections;

public:
	enum {
		NOTIFICATION_POSTINITIALIZE = 0,
		NOTIFICATION_PREDELETE = 1,
	};
	virtual ~Object() {}
	template <typename T>
	static T *cast_to(Object *p) { return dynamic_cast<T *>(p); }
	template <typename T>
	static const T *cast_to(const Object *p) { return dynamic_cast<const T *>(p); }
	void notify_property_list_changed() {}
	// signals: connected callables run synchronously, in connection order, like the engine's default
	template <typename... A>
	int emit_signal(const StringName &p_name, A...) {
		auto it = signal_connections.find(p_name.std_str());
		if (it != signal_connections.end()) {
			std::vector<Callable> targets = it->second;
			for (Callable &c : targets) {
				if (c.fn) {
					c.fn();
				}
			}
		}
		return 0;
	}
	int connect(const StringName &p_name, const Callable &p_callable, uint32_t = 0) {
		signal_connections[p_name.std_str()].push_back(p_callable);
		return 0;
	}
	void disconnect(const StringName &p_name, const Callable &p_callable) {
		std::vector<Callable> &v = signal_connections[p_name.std_str()];
		for (size_t k = 0; k < v.size(); k++) {
			if (v[k] == p_callable) {
				v.erase(v.begin() + (long)k);
				return;
			}
		}
	}
	bool is_connected(const StringName &p_name, const Callable &p_callable) const {
		auto it = signal_connections.find(p_name.std_str());
		if (it == signal_connections.end()) {
			return false;
		}
		for (const Callable &c : it->second) {
			if (c == p_callable) {
				return true;
			}
		}
		return false;
	}
	template <typename... A>
	void call_deferred(const StringName &, A...) {}
	void set_message_translation(bool) {}
	uint64_t get_instance_id() const { return (uint64_t)(uintptr_t)this; }
	virtual String get_class() const { return String("Object"); }
	bool is_class(const String &) const { return false; }
	// Object::set / Object::get: the property path entry points.  GDCLASS chains _setv / _getv from the
	// most derived class's _set / _get down, as the engine's macro does.
	void set(const StringName &p_name, const Variant &p_value, bool *r_valid = nullptr) {
		bool ok = _setv(p_name, p_value);
		if (r_valid) {
			*r_valid = ok;
		}
	}
	inline Variant get(const StringName &p_name, bool *r_valid = nullptr) const;
	typedef bool (Object::*SetMethodShim)(const StringName &, const Variant &);
	typedef bool (Object::*GetMethodShim)(const StringName &, Variant &) const;
	virtual bool _setv(const StringName &, const Variant &) { return false; }
	virtual bool _getv(const StringName &, Variant &) const { return false; }

protected:
	static void _bind_methods() {}
	void _notification(int) {}
	bool _set(const StringName &, const Variant &) { return false; }
	bool _get(const StringName &, Variant &) const { return false; }
	static SetMethodShim _get_set_shim() { return &Object::_set; }
	static GetMethodShim _get_get_shim() { return &Object::_get; }
};

#define GDCLASS(m_class, m_inherits)                                                                   \
public:                                                                                                \
	typedef m_class self_type;                                                                         \
	typedef m_inherits super_type;                                                                     \
	virtual String get_class() const override { return String(#m_class); }                             \
	static String get_class_static() { return String(#m_class); }                                      \
	virtual bool _setv(const StringName &p_name, const Variant &p_value) override {                    \
		if (m_inherits::_setv(p_name, p_value)) {                                                      \
			return true;                                                                               \
		}                                                                                              \
		if (m_class::_get_set_shim() != m_inherits::_get_set_shim()) {                                 \
			return _set(p_name, p_value);                                                              \
		}                                                                                              \
		return false;                                                                                  \
	}                                                                                                  \
	virtual bool _getv(const StringName &p_name, Variant &r_ret) const override {                      \
		if (m_class::_get_get_shim() != m_inherits::_get_get_shim()) {                                 \
			if (_get(p_name, r_ret)) {                                                                 \
				return true;                                                                           \
			}                                                                                          \
		}                                                                                              \
		return m_inherits::_getv(p_name, r_ret);                                                       \
	}                                                                                                  \
                                                                                                       \
protected:                                                                                             \
	static Object::SetMethodShim _get_set_shim() {                                                     \
		return static_cast<Object::SetMethodShim>(&m_class::_set);                                     \
	}                                                                                                  \
	static Object::GetMethodShim _get_get_shim() {                                                     \
		return static_cast<Object::GetMethodShim>(&m_class::_get);                                     \
	}                                                                                                  \
                                                                                                       \
private:

class RefCounted : public Object {
	GDCLASS(RefCounted, Object);
	int refcount = 0;

public:
	void reference() { refcount++; }
	bool unreference() { return --refcount <= 0; }
	int get_reference_count() const { return refcount; }
	bool init_ref() {
		refcount++;
		return true;
	}
};

template <typename T>
class Ref {
	T *p = nullptr;
	void ref_pointer(T *n) {
		if (n) {
			n->reference();
		}
		T *old = p;
		p = n;
		if (old && old->unreference()) {
			delete old;
		}
	}

public:
	Ref() {}
	Ref(T *n) { ref_pointer(n); }
	Ref(const Ref &o) { ref_pointer(o.p); }
	template <typename U>
	Ref(const Ref<U> &o) { ref_pointer(dynamic_cast<T *>(const_cast<U *>(o.ptr()))); }
	inline Ref(const Variant &v);
	~Ref() { unref(); }
	Ref &operator=(const Ref &o) {
		ref_pointer(o.p);
		return *this;
	}
	template <typename U>
	Ref &operator=(const Ref<U> &o) {
		ref_pointer(dynamic_cast<T *>(const_cast<U *>(o.ptr())));
		return *this;
	}
	inline Ref &operator=(const Variant &v);
	void unref() { ref_pointer(nullptr); }
	void instantiate() { ref_pointer(new T); }
	template <typename... A>
	void instantiate(A... a) { ref_pointer(new T(a...)); }
	bool is_valid() const { return p != nullptr; }
	bool is_null() const { return p == nullptr; }
	T *ptr() const { return p; }
	T *operator->() const { return p; }
	T *operator*() const { return p; }
	bool operator==(const Ref &o) const { return p == o.p; }
	bool operator!=(const Ref &o) const { return p != o.p; }
	bool operator==(const T *o) const { return p == o; }
	bool operator!=(const T *o) const { return p != o; }
	bool operator<(const Ref &o) const { return p < o.p; }
};

class Resource : public RefCounted {
	GDCLASS(Resource, RefCounted);
	String name;

public:
	void set_name(const String &p) { name = p; }
	String get_name() const { return name; }
	void emit_changed() {}
	virtual void set_path(const String &, bool = false) {}
};

// ---------------------------------------------------------------- Variant (only the kinds the module stores)
class Variant {
public:
	enum Type {
		NIL,
		BOOL,
		INT,
		FLOAT,
		STRING,
		VECTOR2,
		VECTOR3,
		VECTOR4,
		QUATERNION,
		BASIS,
		TRANSFORM3D,
		COLOR,
		STRING_NAME,
		NODE_PATH,
		OBJECT,
		DICTIONARY,
		ARRAY,
		PACKED_VECTOR3_ARRAY,
		VARIANT_MAX
	};

private:
	Type type = NIL;
	bool b = false;
	int64_t i = 0;
	double f = 0;
	String s;
	Vector4 v4;
	Transform3D xf;
	Object *obj = nullptr; // non-owning for plain objects
	Ref<RefCounted> ref; // owning when the object is reference counted

public:
	Variant() {}
	Variant(bool p) : type(BOOL), b(p) {}
	Variant(int p) : type(INT), i(p) {}
	Variant(unsigned p) : type(INT), i(p) {}
	Variant(int64_t p) : type(INT), i(p) {}
	Variant(uint64_t p) : type(INT), i((int64_t)p) {}
	Variant(float p) : type(FLOAT), f(p) {}
	Variant(double p) : type(FLOAT), f(p) {}
	Variant(const char *p) : type(STRING), s(p) {}
	Variant(const String &p) : type(STRING), s(p) {}
	Variant(const StringName &p) : type(STRING_NAME), s(p) {}
	Variant(const NodePath &p) : type(NODE_PATH), s(p.str()) {}
	Variant(const Vector2 &p) : type(VECTOR2), v4(p.x, p.y, 0, 0) {}
	Variant(const Vector3 &p) : type(VECTOR3), v4(p.x, p.y, p.z, 0) {}
	Variant(const Vector4 &p) : type(VECTOR4), v4(p) {}
	Variant(const Quaternion &p) : type(QUATERNION), v4(p.x, p.y, p.z, p.w) {}
	Variant(const Basis &p) : type(BASIS) { xf.basis = p; }
	Variant(const Transform3D &p) : type(TRANSFORM3D), xf(p) {}
	Variant(Object *p) : type(OBJECT), obj(p) {
		if (RefCounted *r = dynamic_cast<RefCounted *>(p)) {
			ref = Ref<RefCounted>(r);
		}
	}
	template <typename T>
	Variant(const Ref<T> &p) : type(p.is_valid() ? OBJECT : NIL), obj(p.ptr()) {
		if (p.is_valid()) {
			ref = Ref<RefCounted>(static_cast<RefCounted *>(p.ptr()));
		}
	}
	Type get_type() const { return type; }
	bool is_null() const { return type == NIL || (type == OBJECT && obj == nullptr); }
	Object *get_validated_object() const { return obj; }
	operator bool() const { return type == BOOL ? b : (type == INT ? i != 0 : (type == FLOAT ? f != 0 : false)); }
	operator int() const { return (int)(int64_t) * this; }
	operator unsigned() const { return (unsigned)(int64_t) * this; }
	operator int64_t() const { return type == INT ? i : (type == FLOAT ? (int64_t)f : (type == BOOL ? (int64_t)b : 0)); }
	operator float() const { return (float)(double)*this; }
	operator double() const { return type == FLOAT ? f : (type == INT ? (double)i : (type == BOOL ? (double)b : 0.0)); }
	operator String() const { return s; }
	operator StringName() const { return StringName(s); }
	operator NodePath() const { return NodePath(s); }
	operator Vector2() const { return Vector2(v4.x, v4.y); }
	operator Vector3() const { return Vector3(v4.x, v4.y, v4.z); }
	operator Vector4() const { return v4; }
	operator Quaternion() const { return Quaternion(v4.x, v4.y, v4.z, v4.w); }
	operator Basis() const { return xf.basis; }
	operator Transform3D() const { return xf; }
	operator Object *() const { return obj; }
	bool operator==(const Variant &o) const {
		return type == o.type && b == o.b && i == o.i && f == o.f && s == o.s && v4 == o.v4 && xf == o.xf && obj == o.obj;
	}
	bool operator!=(const Variant &o) const { return !(*this == o); }
};

inline Variant Object::get(const StringName &p_name, bool *r_valid) const {
	Variant r;
	bool ok = _getv(p_name, r);
	if (r_valid) {
		*r_valid = ok;
	}
	return r;
}

template <typename T>
inline Ref<T>::Ref(const Variant &v) {
	ref_pointer(dynamic_cast<T *>(v.get_validated_object()));
}
template <typename T>
inline Ref<T> &Ref<T>::operator=(const Variant &v) {
	ref_pointer(dynamic_cast<T *>(v.get_validated_object()));
	return *this;
}

class Array {
	std::vector<Variant> a;

public:
	int size() const { return (int)a.size(); }
	void push_back(const Variant &v) { a.push_back(v); }
	void append(const Variant &v) { a.push_back(v); }
	void resize(int n) { a.resize((size_t)n); }
	void clear() { a.clear(); }
	Variant &operator[](int i) { return a[(size_t)i]; }
	const Variant &operator[](int i) const { return a[(size_t)i]; }
};
template <typename T>
class TypedArray : public Array {
public:
	TypedArray() {}
	TypedArray(const Array &o) : Array(o) {}
};
class Dictionary {
	std::vector<std::pair<Variant, Variant>> d;

public:
	Variant &operator[](const Variant &k) {
		for (auto &e : d) {
			if (e.first == k) {
				return e.second;
			}
		}
		d.push_back({ k, Variant() });
		return d.back().second;
	}
	bool has(const Variant &k) const {
		for (auto &e : d) {
			if (e.first == k) {
				return true;
			}
		}
		return false;
	}
	int size() const { return (int)d.size(); }
};

// a weak reference that does not keep its target alive (engine: WeakRef stores an ObjectID)
class WeakRef : public RefCounted {
	GDCLASS(WeakRef, RefCounted);
	RefCounted *target = nullptr;

public:
	Variant get_ref() const { return target ? Variant(Ref<RefCounted>(target)) : Variant(); }
	template <typename T>
	void set_ref(const Ref<T> &p) { target = p.ptr(); }
	void set_obj(Object *p) { target = dynamic_cast<RefCounted *>(p); }
};

struct PropertyInfo {
	Variant::Type type = Variant::NIL;
	String name;
	PropertyHint hint = PROPERTY_HINT_NONE;
	String hint_string;
	uint32_t usage = PROPERTY_USAGE_DEFAULT;
	StringName class_name;
	PropertyInfo() {}
	PropertyInfo(Variant::Type p_type, const String &p_name, PropertyHint p_hint = PROPERTY_HINT_NONE, const String &p_hint_string = "",
			uint32_t p_usage = PROPERTY_USAGE_DEFAULT, const StringName &p_class_name = StringName()) :
			type(p_type), name(p_name), hint(p_hint), hint_string(p_hint_string), usage(p_usage), class_name(p_class_name) {}
};
struct MethodInfo {
	String name;
	template <typename... A>
	MethodInfo(const String &p_name, A...) : name(p_name) {}
};

// ---------------------------------------------------------------- ClassDB: binding is inert
struct MethodDefinitionShim {};
template <typename... A>
inline MethodDefinitionShim D_METHOD(const char *, A...) { return MethodDefinitionShim(); }
#define DEFVAL(m_v) (m_v)
class ClassDB {
public:
	template <typename M, typename... A>
	static void bind_method(MethodDefinitionShim, M, A...) {}
	template <typename M, typename... A>
	static void bind_static_method(const StringName &, MethodDefinitionShim, M, A...) {}
	template <typename T>
	static void register_class() {}
	static bool class_exists(const StringName &) { return false; }
};
#define ADD_PROPERTY(m_property, m_setter, m_getter) ((void)0)
#define ADD_PROPERTYI(m_property, m_setter, m_getter, m_index) ((void)0)
#define ADD_GROUP(m_name, m_prefix) ((void)0)
#define ADD_SUBGROUP(m_name, m_prefix) ((void)0)
#define ADD_SIGNAL(m_signal) ((void)0)
#define ADD_ARRAY_COUNT(m_label, m_count_property, m_count_property_setter, m_count_property_getter, m_prefix) ((void)0)
#define BIND_ENUM_CONSTANT(m_c) ((void)0)
#define BIND_CONSTANT(m_c) ((void)0)
#define GDREGISTER_CLASS(m_class) ((void)0)
#define VARIANT_ENUM_CAST(m_enum)

class Engine {
public:
	static Engine *get_singleton() {
		static Engine e;
		return &e;
	}
	bool is_editor_hint() const { return false; }
};

// ---------------------------------------------------------------- scene stand-ins
class SceneTree;
class Node : public Object {
	GDCLASS(Node, Object);

protected:
	Node *parent_node = nullptr;
	std::vector<Node *> child_nodes;
	StringName node_name;
	bool inside_tree = true;

public:
	enum {
		NOTIFICATION_ENTER_TREE = 10,
		NOTIFICATION_EXIT_TREE = 11,
		NOTIFICATION_READY = 13,
		NOTIFICATION_PROCESS = 17,
		NOTIFICATION_INTERNAL_PROCESS = 25,
	};
	// node paths: "Name" or "../Name" resolve among this node's (or its parent's) children; enough for the
	// target-node lookups the solver does (src/many_bone_ik_3d.cpp, src/ik_effector_3d.cpp)
	Node *get_node_or_null(const NodePath &p_path) const {
		const Node *cur = this;
		String p = p_path;
		int n = p.get_slice_count("/");
		if (p.is_empty()) {
			return nullptr;
		}
		for (int k = 0; k < n && cur; k++) {
			String part = p.get_slice("/", k);
			if (part == "..") {
				cur = cur->parent_node;
			} else if (part == "." || part.is_empty()) {
			} else {
				const Node *found = nullptr;
				for (Node *c : cur->child_nodes) {
					if ((String)c->node_name == part) {
						found = c;
						break;
					}
				}
				cur = found;
			}
		}
		return const_cast<Node *>(cur);
	}
	Node *get_node(const NodePath &p) const { return get_node_or_null(p); }
	bool has_node(const NodePath &p) const { return get_node_or_null(p) != nullptr; }
	Node *get_parent() const { return parent_node; }
	Node *get_owner() const { return nullptr; }
	void add_child(Node *c) {
		c->parent_node = this;
		child_nodes.push_back(c);
	}
	int get_child_count() const { return (int)child_nodes.size(); }
	Node *get_child(int i) const { return child_nodes[(size_t)i]; }
	void set_name(const String &p) { node_name = StringName(p); }
	StringName get_name() const { return node_name; }
	bool is_inside_tree() const { return inside_tree; }
	bool is_node_ready() const { return true; }
	bool is_ancestor_of(const Node *p) const {
		for (const Node *c = p ? p->parent_node : nullptr; c; c = c->parent_node) {
			if (c == this) {
				return true;
			}
		}
		return false;
	}
	NodePath get_path_to(const Node *) const { return NodePath(); }
	NodePath get_path() const { return NodePath(node_name); }
	SceneTree *get_tree() const { return nullptr; }
	void set_process(bool) {}
	void set_process_internal(bool) {}
	void set_physics_process(bool) {}
	void set_process_priority(int) {}
	void queue_free() {}
	void update_configuration_warnings() {}
	virtual PackedStringArray get_configuration_warnings() const { return PackedStringArray(); }
};

class Node3D : public Node {
	GDCLASS(Node3D, Node);

protected:
	Transform3D xform_local;

public:
	enum {
		NOTIFICATION_TRANSFORM_CHANGED = 2000,
		NOTIFICATION_VISIBILITY_CHANGED = 43,
	};
	// the stand-in scene is flat: a node's global transform is the transform it was given
	Transform3D get_global_transform() const { return xform_local; }
	void set_global_transform(const Transform3D &t) { xform_local = t; }
	Transform3D get_transform() const { return xform_local; }
	void set_transform(const Transform3D &t) { xform_local = t; }
	bool is_visible_in_tree() const { return true; }
	bool is_visible() const { return true; }
	void update_gizmos() {}
	void set_notify_transform(bool) {}
	void set_notify_local_transform(bool) {}
};

class SceneTree : public Object {
	GDCLASS(SceneTree, Object);
};

class Skeleton3D : public Node3D {
	GDCLASS(Skeleton3D, Node3D);
	struct BoneShim {
		String name;
		int parent = -1;
		Transform3D rest;
		Transform3D pose; // what get_bone_pose() returns (the engine's pose_cache, composed from position/rotation/scale)
		bool pose_cache_dirty = false;
		Vector3 pose_position;
		Quaternion pose_rotation;
		Vector3 pose_scale = Vector3(1, 1, 1);
		Vector<int> children;
	};
	std::vector<BoneShim> bones;

public:
	int add_bone(const String &p_name) {
		BoneShim b;
		b.name = p_name;
		bones.push_back(b);
		return (int)bones.size() - 1;
	}
	void set_bone_parent(int p_bone, int p_parent) {
		bones[(size_t)p_bone].parent = p_parent;
		// children are kept in ascending bone index, as Skeleton3D::_update_process_order leaves them
		for (auto &b : bones) {
			b.children.clear();
		}
		for (int i = 0; i < (int)bones.size(); i++) {
			if (bones[(size_t)i].parent >= 0) {
				bones[(size_t)bones[(size_t)i].parent].children.push_back(i);
			}
		}
	}
	int get_bone_count() const { return (int)bones.size(); }
	int find_bone(const String &p_name) const {
		for (int i = 0; i < (int)bones.size(); i++) {
			if (bones[(size_t)i].name == p_name) {
				return i;
			}
		}
		return -1;
	}
	String get_bone_name(int p_bone) const {
		ERR_FAIL_INDEX_V(p_bone, (int)bones.size(), String());
		return bones[(size_t)p_bone].name;
	}
	int get_bone_parent(int p_bone) const {
		ERR_FAIL_INDEX_V(p_bone, (int)bones.size(), -1);
		return bones[(size_t)p_bone].parent;
	}
	Vector<int> get_bone_children(int p_bone) const {
		ERR_FAIL_INDEX_V(p_bone, (int)bones.size(), Vector<int>());
		return bones[(size_t)p_bone].children;
	}
	Vector<int> get_parentless_bones() const {
		Vector<int> r;
		for (int i = 0; i < (int)bones.size(); i++) {
			if (bones[(size_t)i].parent < 0) {
				r.push_back(i);
			}
		}
		return r;
	}
	void set_bone_rest(int p_bone, const Transform3D &t) { bones[(size_t)p_bone].rest = t; }
	Transform3D get_bone_rest(int p_bone) const { return bones[(size_t)p_bone].rest; }
	// get_bone_pose() is the engine's pose cache (scene/3d/skeleton_3d.cpp, Bone::update_pose_cache): once one of the three
	// component setters ran it is recomposed, Transform3D(Basis(pose_rotation, pose_scale), pose_position).  That is the
	// pose the module's next frame re-seeds its IK bones from (src/many_bone_ik_3d.cpp:1084, :91-102, src/ik_bone_3d.cpp:166).
	Transform3D get_bone_pose(int p_bone) const {
		ERR_FAIL_INDEX_V(p_bone, (int)bones.size(), Transform3D());
		BoneShim &b = const_cast<Skeleton3D *>(this)->bones[(size_t)p_bone];
		if (b.pose_cache_dirty) {
			b.pose = Transform3D(Basis(b.pose_rotation, b.pose_scale), b.pose_position);
			b.pose_cache_dirty = false;
		}
		return b.pose;
	}
	// The harness seeds each bone with a raw local Transform3D -- the C ABI's start_pose record, i.e. "what get_bone_pose()
	// returns" -- and the stand-in hands exactly those bits back (the engine would decompose and recompose them; a caller of
	// the C ABI passes poses it read from get_bone_pose(), which already went through that).
	void set_bone_pose(int p_bone, const Transform3D &t) {
		BoneShim &b = bones[(size_t)p_bone];
		b.pose = t;
		b.pose_cache_dirty = false;
		b.pose_position = t.origin;
		b.pose_rotation = t.basis.get_rotation_quaternion();
		b.pose_scale = t.basis.get_scale();
	}
	// the three setters IKBone3D::set_skeleton_bone_pose calls (src/ik_bone_3d.cpp:173-178)
	void set_bone_pose_position(int p_bone, const Vector3 &p) {
		bones[(size_t)p_bone].pose_position = p;
		bones[(size_t)p_bone].pose_cache_dirty = true;
	}
	void set_bone_pose_rotation(int p_bone, const Quaternion &q) {
		bones[(size_t)p_bone].pose_rotation = q;
		bones[(size_t)p_bone].pose_cache_dirty = true;
	}
	void set_bone_pose_scale(int p_bone, const Vector3 &s) {
		bones[(size_t)p_bone].pose_scale = s;
		bones[(size_t)p_bone].pose_cache_dirty = true;
	}
	Vector3 get_bone_pose_position(int p_bone) const { return bones[(size_t)p_bone].pose_position; }
	Quaternion get_bone_pose_rotation(int p_bone) const { return bones[(size_t)p_bone].pose_rotation; }
	Vector3 get_bone_pose_scale(int p_bone) const { return bones[(size_t)p_bone].pose_scale; }
	Transform3D get_bone_global_pose(int p_bone) const {
		ERR_FAIL_INDEX_V(p_bone, (int)bones.size(), Transform3D());
		Transform3D t = get_bone_pose(p_bone);
		int p = bones[(size_t)p_bone].parent;
		return p >= 0 ? get_bone_global_pose(p) * t : t;
	}
	void force_update_all_bone_transforms() {}
	void reset_bone_poses() {
		for (int i = 0; i < (int)bones.size(); i++) {
			set_bone_pose(i, bones[(size_t)i].rest);
		}
	}
};

class SkeletonModifier3D : public Node3D {
	GDCLASS(SkeletonModifier3D, Node3D);
	Skeleton3D *skeleton_shim = nullptr;

protected:
	virtual void _process_modification() {}
	virtual void _skeleton_changed(Skeleton3D *, Skeleton3D *) {}
	void _notification(int) {}
	static void _bind_methods() {}

public:
	Skeleton3D *get_skeleton() const { return skeleton_shim; }
	// engine: the modifier's parent Skeleton3D; here it is attached explicitly
	void shim_attach_skeleton(Skeleton3D *p_skeleton) {
		Skeleton3D *old = skeleton_shim;
		skeleton_shim = p_skeleton;
		if (p_skeleton) {
			p_skeleton->add_child(this);
		}
		_skeleton_changed(old, p_skeleton);
	}
	// engine: SkeletonModifier3D::process_modification() -> virtual _process_modification(), then the
	// "modification_processed" signal
	void process_modification() {
		_process_modification();
		emit_signal(SNAME("modification_processed"));
	}
	void set_active(bool) {}
	bool is_active() const { return true; }
	bool is_enabled() const { return true; } // called at src/many_bone_ik_3d.cpp:679
	void set_influence(real_t) {}
	real_t get_influence() const { return 1; }
};
