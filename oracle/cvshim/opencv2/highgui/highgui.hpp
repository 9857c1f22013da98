// ORACLE: see cvshim.hpp
#pragma once
#include "../../cvshim.hpp"
