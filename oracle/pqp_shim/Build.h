// TEST INFRASTRUCTURE (oracle) -- PQP's Build.h is included by the reference
// (C2A/src/C2A.cpp:13) but nothing from it is called; C2A has its own builder.
#ifndef PQP_SHIM_BUILD_H
#define PQP_SHIM_BUILD_H
#include "PQP.h"
#endif
