#include "streamgen.h"
namespace sg { std::vector<uint8_t> write_mpeg2(const Mpeg2Opts &) { return {}; } }
