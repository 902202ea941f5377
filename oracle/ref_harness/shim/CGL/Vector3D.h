// shim: src/marching.h:2 includes "CGL/Vector3D.h" but the file on disk is vector3D.h (case-sensitive FS)
#include "CGL/vector3D.h"
