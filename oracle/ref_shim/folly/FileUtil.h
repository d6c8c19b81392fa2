// ORACLE shim: see dynamic.h
#pragma once
#include "dynamic.h"
