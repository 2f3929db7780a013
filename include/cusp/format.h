// cusp/format.h — format tags live in cusp/memory.h in this shim
#pragma once
#include "memory.h"
