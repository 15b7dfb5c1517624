#pragma once
#define UCOSLAM_API
