// ORACLE — TEST INFRASTRUCTURE ONLY.  Runs every known-answer test of the restatement.
#include <cstdio>

#include "kat_check.h"

int runDeformKats();
int runCompressionKats();
int runDeltaKats();

int main() {
  runDeformKats();
  const int after_deform = kat::failures();
  runCompressionKats();
  const int after_compression = kat::failures();
  runDeltaKats();
  std::printf("oracle KAT: %d checks, %d failures (%d deformation, %d compression, %d delta compression)\n",
              kat::checks(), kat::failures(), after_deform, after_compression - after_deform,
              kat::failures() - after_compression);
  return kat::failures() ? 1 : 0;
}
