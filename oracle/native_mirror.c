#include "native_mirror.h"
