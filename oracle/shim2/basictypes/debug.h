#pragma once   // TEST INFRASTRUCTURE ONLY: the reference's debug macros (src/basictypes/debug.h) as no-ops
#define _debug_msg(x, level)
#define _debug_msg_(x)
#define _debug_exec(level, x)
#define _debug_exec_(x)
namespace ucoslam { namespace debug { struct Debug { static int getLevel() { return 0; } }; } }
