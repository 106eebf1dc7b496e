// stand-in for <pcl/filters/voxel_grid.h>: included by the reference node (src/processor.cpp:48), never used. TEST INFRASTRUCTURE ONLY.
#pragma once
