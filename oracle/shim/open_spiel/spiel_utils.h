// TEST INFRASTRUCTURE ONLY (oracle/): see spiel.h in this directory.
// The reference includes this header (twixt.cc:25) but everything it uses from
// it is supplied by the shim's spiel.h.
#ifndef ORACLE_SHIM_OPEN_SPIEL_SPIEL_UTILS_H_
#define ORACLE_SHIM_OPEN_SPIEL_SPIEL_UTILS_H_
#include "open_spiel/spiel.h"
#endif  // ORACLE_SHIM_OPEN_SPIEL_SPIEL_UTILS_H_
