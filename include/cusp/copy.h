// cusp/copy.h — cusp::copy(src, dst): same format, any memory spaces (defined in convert.h)
#pragma once
#include "convert.h"
