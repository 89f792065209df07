#pragma once
#include "cuda_emul.h"
