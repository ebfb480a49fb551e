#pragma once
#include <grid_map_ros/GridMapRosConverter.hpp>
