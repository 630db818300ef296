// TEST INFRASTRUCTURE ONLY (oracle/): minimal stand-in for Xilinx's ap_int.h.
// ap_uint<W>/ap_int<W> with W<=31: value is stored wrapped to W bits and converts implicitly to
// int, so mixed expressions follow ordinary C integer arithmetic (a superset of what the
// reference designs rely on: loop counters, indices, comparisons, stream payloads).
#ifndef ORACLE_SHIM_AP_INT_H
#define ORACLE_SHIM_AP_INT_H
#include <iostream>
template <int W> struct ap_int;
template <int W>
struct ap_uint {
    static_assert(W >= 1 && W <= 31, "shim supports 1..31 bits");
    unsigned v;
    static unsigned wrap(long long x) { return (unsigned)(x & ((1LL << W) - 1)); }
    ap_uint() : v(0) {}
    ap_uint(int x) : v(wrap(x)) {}
    ap_uint(unsigned x) : v(wrap(x)) {}
    ap_uint(long x) : v(wrap(x)) {}
    ap_uint(long long x) : v(wrap(x)) {}
    ap_uint(unsigned long x) : v(wrap((long long)x)) {}
    template <int W2> ap_uint(const ap_uint<W2>& o) : v(wrap(o.v)) {}
    template <int W2> ap_uint(const ap_int<W2>& o);
    operator int() const { return (int)v; }
    ap_uint& operator++() { v = wrap((long long)v + 1); return *this; }
    ap_uint operator++(int) { ap_uint t = *this; ++*this; return t; }
    ap_uint& operator--() { v = wrap((long long)v - 1); return *this; }
    ap_uint operator--(int) { ap_uint t = *this; --*this; return t; }
    ap_uint& operator+=(int x) { v = wrap((long long)v + x); return *this; }
    ap_uint& operator-=(int x) { v = wrap((long long)v - x); return *this; }
};
template <int W>
struct ap_int {
    static_assert(W >= 2 && W <= 31, "shim supports 2..31 bits");
    int v;
    static int wrap(long long x) {
        long long m = x & ((1LL << W) - 1);
        if (m & (1LL << (W - 1))) m -= (1LL << W);
        return (int)m;
    }
    ap_int() : v(0) {}
    ap_int(int x) : v(wrap(x)) {}
    ap_int(unsigned x) : v(wrap(x)) {}
    ap_int(long x) : v(wrap(x)) {}
    ap_int(long long x) : v(wrap(x)) {}
    template <int W2> ap_int(const ap_uint<W2>& o) : v(wrap(o.v)) {}
    operator int() const { return v; }
    ap_int& operator++() { v = wrap((long long)v + 1); return *this; }
    ap_int operator++(int) { ap_int t = *this; ++*this; return t; }
    ap_int& operator--() { v = wrap((long long)v - 1); return *this; }
    ap_int operator--(int) { ap_int t = *this; --*this; return t; }
};
template <int W> template <int W2> ap_uint<W>::ap_uint(const ap_int<W2>& o) : v(wrap(o.v)) {}
#endif
