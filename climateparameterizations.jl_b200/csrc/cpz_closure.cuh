#pragma once
#include "cpz_device.cuh"
