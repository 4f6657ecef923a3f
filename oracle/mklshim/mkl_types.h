/* MKL-compat shim (test infrastructure): see mkl.h */
#include "mkl.h"
