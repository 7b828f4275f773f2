// Part of the from-scratch GLM stand-in (see detail_fgoicp.hpp for scope and rationale).
#pragma once
#include "detail_fgoicp.hpp"
