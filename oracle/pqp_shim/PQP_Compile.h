// TEST INFRASTRUCTURE (oracle) -- clean-room subset of PQP's PQP_Compile.h.
// PQP_REAL must be double (the reference passes double[3][3] where PQP_REAL[3][3]
// is expected, C2A/src/C2A.cpp:1118-1119) and both BV types must be on
// (->Tr used unguarded at C2A/src/C2A.cpp:1205; C2A/src/C2A_PQP.cpp:298-302 only
// compiles in the OBB branch).
#ifndef PQP_SHIM_COMPILE_H
#define PQP_SHIM_COMPILE_H

typedef double PQP_REAL;

#define RSS_TYPE 1
#define OBB_TYPE 2
#define PQP_BV_TYPE (RSS_TYPE | OBB_TYPE)

#endif
