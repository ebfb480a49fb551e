#pragma once
#include <stdint.h>
#include <string>
namespace ros {
struct Time { uint32_t sec, nsec; Time() : sec(0), nsec(0) {} };
class Publisher {};
class NodeHandle {};
}  // namespace ros
