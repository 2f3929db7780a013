// cusp/blas.h — forwards to cusp/blas/blas.h (the reference keeps both spellings)
#pragma once
#include "blas/blas.h"
