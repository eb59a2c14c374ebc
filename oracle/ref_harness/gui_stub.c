/* TEST INFRASTRUCTURE -- not product code.
 * Headless replacement for the reference's ncurses TUI (gui.c). Implements the
 * 14 entry points declared in /root/reference/gui.h:64-77 as no-ops; status
 * lines go to stderr when ORACLE_VERBOSE is set. Needed because ncurses is not
 * installed here and the oracle must run without a terminal. */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include "gui.h"

static int verbose(void) {
    static int v = -1;
    if (v < 0) v = getenv("ORACLE_VERBOSE") != NULL;
    return v;
}

void gui_init(void) {}
int gui_getch(void) { return -1; }
void gui_destroy(void) {}

void gui_mvwprintw(window_panel_t w, int y, int x, const char *fmt, ...) {
    (void) y; (void) x;
    if (!verbose()) return;
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[panel %d] ", (int) w);
    vfprintf(stderr, fmt, ap);
    fputc('\n', stderr);
    va_end(ap);
}

void gui_status_wprintw(status_color_t clr, const char *fmt, ...) {
    if (!verbose() && clr != RED) return;
    va_list ap;
    va_start(ap, fmt);
    vfprintf(stderr, fmt, ap);
    va_end(ap);
}

void gui_colorpair(window_panel_t w, unsigned clr, attr_status_t onoff) { (void) w; (void) clr; (void) onoff; }
void gui_top_panel(window_panel_t p) { (void) p; }
void gui_toggle_current_panel(void) {}
void gui_show_panel(window_panel_t p, attr_status_t onoff) { (void) p; (void) onoff; }
void gui_show_speed(float speed) { (void) speed; }
void gui_show_heading(float hdg) { (void) hdg; }
void gui_show_vertical_speed(float vs) { (void) vs; }
void gui_show_location(void *l) { (void) l; }
void gui_show_target(void *t) { (void) t; }
