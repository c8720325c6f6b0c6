#pragma once
#include "grid.cuh"
