#pragma once   // TEST INFRASTRUCTURE ONLY: the reference's timer macros (src/basictypes/timers.h) as no-ops
#define __UCOSLAM_ADDTIMER__
#define __UCOSLAM_TIMER_EVENT__(Y)
