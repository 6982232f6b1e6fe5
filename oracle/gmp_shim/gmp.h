/* Minimal GMP declaration shim -- TEST INFRASTRUCTURE ONLY.
 *
 * This image ships the GMP runtime (libgmp.so.10) but not its development
 * header.  The reference's neogb sources only need a couple of dozen mpz_*
 * entry points (for the QQ code paths that are NOT on the GF(p) hot path), so
 * this header declares exactly those against the stable libgmp ABI
 * (__gmpz_* symbols, __mpz_struct = {int alloc; int size; limb*}).
 * It is used only to compile the unmodified reference into oracle/_ref/.
 */
#ifndef F4LA_ORACLE_GMP_SHIM_H
#define F4LA_ORACLE_GMP_SHIM_H

#include <stddef.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long int mp_limb_t;

typedef struct {
    int _mp_alloc;
    int _mp_size;
    mp_limb_t *_mp_d;
} __mpz_struct;

typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

/* lifetime */
void __gmpz_init(mpz_ptr);
void __gmpz_inits(mpz_ptr, ...);
void __gmpz_clear(mpz_ptr);
void __gmpz_clears(mpz_ptr, ...);
void __gmpz_init_set(mpz_ptr, mpz_srcptr);
void __gmpz_init_set_si(mpz_ptr, long);
/* assignment */
void __gmpz_set(mpz_ptr, mpz_srcptr);
void __gmpz_set_si(mpz_ptr, long);
void __gmpz_set_ui(mpz_ptr, unsigned long);
void __gmpz_swap(mpz_ptr, mpz_ptr);
unsigned long __gmpz_get_ui(mpz_srcptr);
/* arithmetic */
void __gmpz_neg(mpz_ptr, mpz_srcptr);
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_submul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_divexact(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_lcm(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_gcd(mpz_ptr, mpz_srcptr, mpz_srcptr);
unsigned long __gmpz_fdiv_ui(mpz_srcptr, unsigned long);
void __gmpz_nextprime(mpz_ptr, mpz_srcptr);
/* predicates */
int __gmpz_cmp_si(mpz_srcptr, long);
int __gmpz_divisible_p(mpz_srcptr, mpz_srcptr);
int __gmpz_divisible_ui_p(mpz_srcptr, unsigned long);
/* formatted output */
int __gmp_printf(const char *, ...);

#define mpz_init            __gmpz_init
#define mpz_inits           __gmpz_inits
#define mpz_clear           __gmpz_clear
#define mpz_clears          __gmpz_clears
#define mpz_init_set        __gmpz_init_set
#define mpz_init_set_si     __gmpz_init_set_si
#define mpz_set             __gmpz_set
#define mpz_set_si          __gmpz_set_si
#define mpz_set_ui          __gmpz_set_ui
#define mpz_swap            __gmpz_swap
#define mpz_get_ui          __gmpz_get_ui
#define mpz_neg             __gmpz_neg
#define mpz_mul             __gmpz_mul
#define mpz_submul          __gmpz_submul
#define mpz_divexact        __gmpz_divexact
#define mpz_lcm             __gmpz_lcm
#define mpz_gcd             __gmpz_gcd
#define mpz_fdiv_ui         __gmpz_fdiv_ui
#define mpz_nextprime       __gmpz_nextprime
#define mpz_cmp_si          __gmpz_cmp_si
#define mpz_divisible_p     __gmpz_divisible_p
#define mpz_divisible_ui_p  __gmpz_divisible_ui_p
#define gmp_printf          __gmp_printf

/* in real gmp.h these two are macros on the struct, not functions */
#define mpz_sgn(Z) ((Z)->_mp_size < 0 ? -1 : ((Z)->_mp_size > 0))

#ifdef __cplusplus
}
#endif
#endif
