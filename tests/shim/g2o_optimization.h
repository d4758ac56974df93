// tests/shim/g2o_optimization.h — forwards to the stand-in declarations (the reference's own header is used in its tree).
#pragma once
#include "urmvo_reference_standin.h"
