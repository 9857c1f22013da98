#pragma once
#include "Frame.h"
