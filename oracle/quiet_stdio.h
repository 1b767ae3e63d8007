// Forced-include for the oracle/_ref build of the UNMODIFIED reference sources (oracle/Makefile).
// TEST INFRASTRUCTURE, not a product component.
//
//  * <vector> and <cstdio>: the reference relies on transitive includes that libstdc++ 13 no
//    longer provides (Threading.h:36 uses std::vector; Chisel.h:62 etc. use printf/puts).
//  * The reference prints several lines per frame and per thread (Chisel.h:62,109,117,126,
//    144-146,158; ChunkManager.cpp:148,675-677). They are not part of the algorithm; they are
//    routed to no-ops so that a pytest run is readable and the timed CPU baseline is not
//    charged for terminal I/O.
#ifndef CVIDS_ORACLE_QUIET_STDIO_H
#define CVIDS_ORACLE_QUIET_STDIO_H
#include <cstdio>
#include <vector>
#include <cmath>
#include <cstdlib>
#include <string>
#include <iostream>
#include <fstream>
static inline int cvids_quiet_printf(const char *, ...) { return 0; }
static inline int cvids_quiet_puts(const char *) { return 0; }
#define printf cvids_quiet_printf
#define puts cvids_quiet_puts
#endif
