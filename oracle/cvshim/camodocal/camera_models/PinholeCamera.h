/* TEST INFRASTRUCTURE ONLY -- see Camera.h next to this file. */
#include "Camera.h"
