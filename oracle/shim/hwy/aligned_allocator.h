#pragma once
#include "hwy/highway.h"
