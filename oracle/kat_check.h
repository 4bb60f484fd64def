// ORACLE — TEST INFRASTRUCTURE ONLY.  Minimal check macros for the known-answer tests.
#pragma once
#include <cmath>
#include <cstdio>
#include <iostream>

namespace kat {
inline int& failures() {
  static int n = 0;
  return n;
}
inline int& checks() {
  static int n = 0;
  return n;
}
}  // namespace kat

#define KAT_EQ(a, b)                                                                  \
  do {                                                                                \
    ++kat::checks();                                                                  \
    if (!((a) == (b))) {                                                              \
      ++kat::failures();                                                              \
      std::cerr << __FILE__ << ":" << __LINE__ << " KAT_EQ(" #a ", " #b ") failed: "   \
                << (a) << " vs " << (b) << "\n";                                      \
    }                                                                                 \
  } while (0)

#define KAT_TRUE(a)                                                                   \
  do {                                                                                \
    ++kat::checks();                                                                  \
    if (!(a)) {                                                                       \
      ++kat::failures();                                                              \
      std::cerr << __FILE__ << ":" << __LINE__ << " KAT_TRUE(" #a ") failed\n";        \
    }                                                                                 \
  } while (0)

#define KAT_NEAR(a, b, tol)                                                           \
  do {                                                                                \
    ++kat::checks();                                                                  \
    if (!(std::fabs(static_cast<double>(a) - static_cast<double>(b)) <= (tol))) {     \
      ++kat::failures();                                                              \
      std::cerr << __FILE__ << ":" << __LINE__ << " KAT_NEAR(" #a ", " #b ") failed: " \
                << (a) << " vs " << (b) << "\n";                                      \
    }                                                                                 \
  } while (0)
