/*
 * shell_oracle_q9.c — the plain-C restatement (shell_oracle.c) compiled for the 9-node
 * MITC9 shell (TACSQuad9Shell / TACSQuad9NonlinearShell, TACSShellElementDefs.h:16-37):
 * TACSShellQuadBasis<3> (TACSShellElementQuadBasis.h:118-735), TACSQuadQuadraticQuadrature
 * (TACSShellElementQuadrature.h:62-110).  TEST INFRASTRUCTURE ONLY (see shell_oracle.h).
 * Entry points: oracle9_* with 27 coordinates, 54 variables and 54 x 54 matrices per element.
 */
#define ORACLE_ORDER 3
#include "shell_oracle.c"
