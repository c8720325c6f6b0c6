// knn.cuh — placeholder, filled in below.
#pragma once
#include "search.cuh"
