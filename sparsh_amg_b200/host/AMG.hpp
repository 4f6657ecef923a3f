// AMG.hpp — drop-in include name of the reference (include/AMG.hpp); everything lives in sparsh_amg.hpp
#include "sparsh_amg.hpp"
