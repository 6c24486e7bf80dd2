// stand-in: the cv types that the reference headers mention live in ref_shim/DBoW3/DBoW3.h (TEST INFRASTRUCTURE)
#include <DBoW3/DBoW3.h>
