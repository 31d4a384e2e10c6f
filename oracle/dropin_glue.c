/* oracle/dropin_glue.c -- TEST INFRASTRUCTURE.  Symbols that lived in the four replaced reference objects and
 * that OTHER reference objects still reference (sgrep.c owned them; newmgrep.c / agrep.c only read them).
 * Nothing here is on the scan path. */
unsigned char TR[256];
