// Shim standing in for abseil's int128.h so that the reference's src/kmer.cpp
// (which only needs a 128-bit unsigned integer with the usual operators)
// compiles without abseil.  Test infrastructure only -- see oracle/README.md.
#pragma once
#include <cstdint>
#include <cstring>
#include <cerrno>
#include <algorithm>
#include <vector>
#include <string>
namespace absl {
typedef unsigned __int128 uint128;
inline uint128 MakeUint128(uint64_t hi, uint64_t lo) { return ((uint128)hi << 64) | (uint128)lo; }
inline uint64_t Uint128Low64(uint128 v) { return (uint64_t)v; }
inline uint64_t Uint128High64(uint128 v) { return (uint64_t)(v >> 64); }
}  // namespace absl
