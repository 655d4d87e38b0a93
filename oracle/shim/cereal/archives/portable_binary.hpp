// TEST INFRASTRUCTURE — not product code.
//
// Stand-in for USCiLab/cereal 1.3.2 PortableBinary{Output,Input}Archive (pinned by the
// reference at libs/CMakeLists.txt:78-96; not vendored, not available offline), just large
// enough for the reference's call sites (src/sdf/SdfFunction.cpp:17,52; OctreeSdf.h:222-238;
// ExactOctreeSdf.h:138-148; TriangleUtils.h:50-54; Mesh.h:65-69; UsefullSerializations.h).
// Published wire behaviour restated here (little-endian host only):
//   * archive ctor writes/reads one byte: 1 = little endian
//   * arithmetic types and enums (as their underlying type) are raw little-endian bytes
//   * std::vector<T>: u64 element count, then the elements (bulk for arithmetic T — same bytes)
//   * std::array<T,N>: the N elements, no count
//   * class types: member save()/load(), else member serialize(), else free serialize()
// No .bin fixture exists in the reference, so the byte layout is "parity unpinned".
#pragma once
#include <cstdint>
#include <iostream>
#include <vector>
#include <array>
#include <type_traits>
#include <memory>
#include <functional>
#include <string>
#include <map>
#include <utility>

namespace cereal {

namespace shim_detail {
template <class T, class A>
auto has_save(int) -> decltype(std::declval<const T&>().save(std::declval<A&>()), std::true_type{});
template <class, class> std::false_type has_save(...);
template <class T, class A>
auto has_load(int) -> decltype(std::declval<T&>().load(std::declval<A&>()), std::true_type{});
template <class, class> std::false_type has_load(...);
template <class T, class A>
auto has_member_serialize(int) -> decltype(std::declval<T&>().serialize(std::declval<A&>()), std::true_type{});
template <class, class> std::false_type has_member_serialize(...);
}  // namespace shim_detail

class PortableBinaryOutputArchive {
    std::ostream& os;
    void raw(const void* p, size_t n) { os.write(static_cast<const char*>(p), std::streamsize(n)); }

    template <class T> void put(const T& t, std::true_type /*arith*/) { raw(&t, sizeof(T)); }
    template <class T> void put(const T& t, std::false_type) { putNonArith(t, std::is_enum<T>{}); }
    template <class T> void putNonArith(const T& t, std::true_type /*enum*/) {
        auto v = static_cast<typename std::underlying_type<T>::type>(t);
        raw(&v, sizeof(v));
    }
    template <class T> void putNonArith(const T& t, std::false_type) {
        putClass(t, decltype(shim_detail::has_save<T, PortableBinaryOutputArchive>(0)){});
    }
    template <class T> void putClass(const T& t, std::true_type /*save()*/) { t.save(*this); }
    template <class T> void putClass(const T& t, std::false_type) {
        putSer(const_cast<T&>(t), decltype(shim_detail::has_member_serialize<T, PortableBinaryOutputArchive>(0)){});
    }
    template <class T> void putSer(T& t, std::true_type) { t.serialize(*this); }
    template <class T> void putSer(T& t, std::false_type) { serialize(*this, t); }  // ADL (glm::serialize)

  public:
    explicit PortableBinaryOutputArchive(std::ostream& s) : os(s) { uint8_t le = 1; raw(&le, 1); }

    template <class... Ts> void operator()(Ts&&... ts) { (one(ts), ...); }

    template <class T> void one(const T& t) { put(t, std::is_arithmetic<T>{}); }
    template <class T> void one(const std::vector<T>& v) {
        uint64_t n = v.size();
        raw(&n, 8);
        for (const auto& e : v) one(e);
    }
    template <class T, size_t N> void one(const std::array<T, N>& v) { for (const auto& e : v) one(e); }
};

class PortableBinaryInputArchive {
    std::istream& is;
    void raw(void* p, size_t n) { is.read(static_cast<char*>(p), std::streamsize(n)); }

    template <class T> void get(T& t, std::true_type) { raw(&t, sizeof(T)); }
    template <class T> void get(T& t, std::false_type) { getNonArith(t, std::is_enum<T>{}); }
    template <class T> void getNonArith(T& t, std::true_type) {
        typename std::underlying_type<T>::type v;
        raw(&v, sizeof(v));
        t = static_cast<T>(v);
    }
    template <class T> void getNonArith(T& t, std::false_type) {
        getClass(t, decltype(shim_detail::has_load<T, PortableBinaryInputArchive>(0)){});
    }
    template <class T> void getClass(T& t, std::true_type) { t.load(*this); }
    template <class T> void getClass(T& t, std::false_type) {
        getSer(t, decltype(shim_detail::has_member_serialize<T, PortableBinaryInputArchive>(0)){});
    }
    template <class T> void getSer(T& t, std::true_type) { t.serialize(*this); }
    template <class T> void getSer(T& t, std::false_type) { serialize(*this, t); }

  public:
    explicit PortableBinaryInputArchive(std::istream& s) : is(s) { uint8_t le = 0; raw(&le, 1); }

    template <class... Ts> void operator()(Ts&&... ts) { (one(ts), ...); }

    template <class T> void one(T& t) { get(t, std::is_arithmetic<T>{}); }
    template <class T> void one(std::vector<T>& v) {
        uint64_t n = 0;
        raw(&n, 8);
        v.resize(n);
        for (auto& e : v) one(e);
    }
    template <class T, size_t N> void one(std::array<T, N>& v) { for (auto& e : v) one(e); }
};

}  // namespace cereal
