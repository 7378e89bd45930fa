// Stand-in for <ros/ros.h>, written for the oracle build only (TEST INFRASTRUCTURE).
// The reference files compiled by oracle/Makefile.ref (distance_heuristics.cpp, data_structs.h) use ROS solely
// for the ROS_ERROR logging macro and rely on <ros/ros.h> pulling in the standard containers.
#pragma once
#include <cmath>
#include <cstdio>
#include <string>
#include <vector>
#define ROS_ERROR(...) do { std::fprintf(stderr, "[ref ROS_ERROR] " __VA_ARGS__); std::fprintf(stderr, "\n"); } while (0)
#define ROS_WARN(...) do { } while (0)
#define ROS_INFO(...) do { } while (0)
