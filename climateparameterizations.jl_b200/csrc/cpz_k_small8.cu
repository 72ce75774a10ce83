#include "cpz_small_impl.h"
namespace cpz {
CPZ_SMALL_DEFINE(8)
}
