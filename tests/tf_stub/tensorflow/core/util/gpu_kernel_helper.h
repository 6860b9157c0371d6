#include "tf_stub_api.h"   // stand-in, see tests/tf_stub/tf_stub_api.h
