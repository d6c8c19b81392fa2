// ORACLE — TEST INFRASTRUCTURE ONLY (see core.hpp): nothing from this module is used on the render path.
#pragma once
#include "core.hpp"
#include "imgcodecs.hpp"
