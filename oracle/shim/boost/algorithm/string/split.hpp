// see ../string.hpp (stand-in; test infrastructure only)
#include "../string.hpp"
