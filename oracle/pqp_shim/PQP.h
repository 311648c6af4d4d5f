// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's PQP.h: the error
// codes and query flags the reference's C2A sources name
// (/root/reference/C2A/src/C2A_PQP.cpp:104,116,139,347,356,875,921).
#ifndef PQP_SHIM_PQP_H
#define PQP_SHIM_PQP_H
#include "PQP_Compile.h"
#include "PQP_Internal.h"

const int PQP_OK = 0;
const int PQP_ERR_MODEL_OUT_OF_MEMORY = -1;
const int PQP_ERR_OUT_OF_MEMORY = -2;
const int PQP_ERR_UNPROCESSED_MODEL = -3;
const int PQP_ERR_BUILD_OUT_OF_SEQUENCE = -4;
const int PQP_ERR_BUILD_EMPTY_MODEL = -5;

const int PQP_ALL_CONTACTS = 1;
const int PQP_FIRST_CONTACT = 2;
#endif
