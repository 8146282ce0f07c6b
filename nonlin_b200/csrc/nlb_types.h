// nlb_types.h — internal alias of the public C ABI types for device code.
#pragma once
#include "../../include/nonlin_batch.h"
