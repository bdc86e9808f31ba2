#pragma once
#include "../marching_cubes.h"
