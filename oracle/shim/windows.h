/*
 * TEST INFRASTRUCTURE ONLY (oracle/): stand-in for <windows.h> so that the
 * reference's Win32-only sources (src/lock.h:19 includes <windows.h>) compile
 * with g++ on Linux, unmodified, from where they lie under /root/reference.
 *
 * What the reference needs from the Win32 headers:
 *   - CRITICAL_SECTION + Initialize/Delete/Enter/Leave  (src/lock.h:25-53)  -> recursive pthread mutex
 *   - InterlockedIncrement/Decrement                    (src/lock.h:59-103, unused class)
 *   - the min()/max() MACROS of minwindef.h             (src/speechPlayer.cpp:36,
 *     src/speechWaveGenerator.cpp:208) -- macro semantics matter: a NaN compares
 *     false, so min(NaN,32000) -> 32000.
 * The standard headers the sources rely on transitively (memcpy, queue, ...)
 * are pulled in BEFORE the macros so the macros cannot break libstdc++.
 */
#ifndef NVSP_ORACLE_FAKE_WINDOWS_H
#define NVSP_ORACLE_FAKE_WINDOWS_H
#include <pthread.h>
#include <cstring>
#include <cstdlib>
#include <cmath>
#include <list>
#include <queue>
#include <iostream>

typedef pthread_mutex_t CRITICAL_SECTION;
typedef long LONG;

static inline void InitializeCriticalSection(CRITICAL_SECTION *cs) {
	pthread_mutexattr_t attr;
	pthread_mutexattr_init(&attr);
	pthread_mutexattr_settype(&attr, PTHREAD_MUTEX_RECURSIVE);
	pthread_mutex_init(cs, &attr);
	pthread_mutexattr_destroy(&attr);
}
static inline void DeleteCriticalSection(CRITICAL_SECTION *cs) { pthread_mutex_destroy(cs); }
static inline void EnterCriticalSection(CRITICAL_SECTION *cs) { pthread_mutex_lock(cs); }
static inline void LeaveCriticalSection(CRITICAL_SECTION *cs) { pthread_mutex_unlock(cs); }
static inline LONG InterlockedIncrement(volatile LONG *v) { return __sync_add_and_fetch(v, 1); }
static inline LONG InterlockedDecrement(volatile LONG *v) { return __sync_sub_and_fetch(v, 1); }

#define max(a, b) (((a) > (b)) ? (a) : (b))
#define min(a, b) (((a) < (b)) ? (a) : (b))
#endif
