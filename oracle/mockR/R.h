/* see Rinternals.h in this directory: mock R API, test infrastructure only */
#ifndef MOCK_R_H
#define MOCK_R_H
#include "Rinternals.h"
#endif
