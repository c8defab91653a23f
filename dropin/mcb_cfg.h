/* Run-time accessors used by the generated config.h (dropin/build_dropin.sh; the
 * oracle/_ref build of the unmodified reference includes the same header) so that
 * ONE binary per (read length, mode) serves every option combination: the
 * reference's wrapper script rewrites config.h and recompiles for each run
 * (minicom:56-103); here the same macros read environment variables instead.
 * `readlen` stays a compile-time constant (std::bitset<2*readlen>,
 * bbhashdict.h:57-63). No behavioural change to the reference. */
#pragma once
#include <stdlib.h>
static inline int mcb_cfg_int(const char *name, int dflt)
{
	const char *s = getenv(name);
	return (s && *s) ? atoi(s) : dflt;
}
static inline const char *mcb_cfg_str(const char *name, const char *dflt)
{
	const char *s = getenv(name);
	return (s && *s) ? s : dflt;
}
