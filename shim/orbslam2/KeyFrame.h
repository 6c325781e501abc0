#include "Frame.h"
