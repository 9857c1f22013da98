// ORACLE: see serialization.hpp
#pragma once
#include "serialization.hpp"
