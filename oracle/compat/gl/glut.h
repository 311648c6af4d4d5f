// TEST INFRASTRUCTURE (oracle): empty stand-in; the reference's library code
// includes GLUT without using it (/root/reference/C2A/src/InterpMotion.cpp:7).
