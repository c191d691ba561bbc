#pragma once
#include <numeric>
namespace thrust {
template <typename It> void sequence(It a, It b) { std::iota(a, b, 0); }
}  // namespace thrust
