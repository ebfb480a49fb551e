#pragma once
#include <string>
#include <vector>
#include <stdint.h>
namespace std_msgs {
struct MultiArrayDimension { std::string label; uint32_t size, stride; MultiArrayDimension() : size(0), stride(0) {} };
struct MultiArrayLayout { std::vector<MultiArrayDimension> dim; uint32_t data_offset; MultiArrayLayout() : data_offset(0) {} };
struct Float32MultiArray { MultiArrayLayout layout; std::vector<float> data; };
}  // namespace std_msgs
