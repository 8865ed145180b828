// stand-in (test infrastructure only)
#include <pcl/point_cloud.h>
