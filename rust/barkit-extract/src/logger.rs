//! The three step lines and the final "Done in ..." line are printed by `bk_run` itself (same text as
//! logger.rs:51-64,105-113 of the reference); this type only keeps the module path alive for callers that
//! construct a Logger around their own loops.
pub struct Logger {
    current: usize,
    total: usize,
    quiet: bool,
}

impl Logger {
    pub fn new(total: usize, quiet: bool) -> Self {
        Logger { current: 0, total, quiet }
    }
    pub fn message(&mut self, text: &str) {
        if self.current < self.total {
            self.current += 1;
            if !self.quiet {
                println!("[{}/{}] {}", self.current, self.total, text);
            }
        } else {
            eprintln!("Warning: Current step exceeds total steps.");
        }
    }
}
