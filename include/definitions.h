// Basic types of the cals:: C++ surface (mirrors reference include/definitions.h:1-20: dim_t, DEBUG, TIME).
#ifndef CALS_B200_DEFINITIONS_H
#define CALS_B200_DEFINITIONS_H

#include <cstddef>

typedef std::size_t dim_t;

#ifdef NDEBUG
#define DEBUG(expr) do { } while (0);
#else
#define DEBUG(expr) expr
#endif

#if WITH_TIME
#define TIME(expr) expr
#else
#define TIME(expr) ;
#endif

#endif
