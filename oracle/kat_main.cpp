// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs every known-answer test of the restatement.
#include <cstdio>

#include "kat_check.h"

int runDeformKats();
int runCompressionKats();

int main() {
  runDeformKats();
  const int after_deform = kat::failures();
  runCompressionKats();
  std::printf("oracle KAT: %d checks, %d failures (%d deformation, %d compression)\n",
              kat::checks(), kat::failures(), after_deform, kat::failures() - after_deform);
  return kat::failures() ? 1 : 0;
}
