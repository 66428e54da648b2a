#!/usr/bin/env perl
# perl/bin/ppp_cutoff_b200.pl -- command-line drop-in for ProPhylo's bin/ppp_cutoff.pl.
#
# Reads a ppp.pl / ppp_b200.pl table (file argument or stdin) and splits it at the slope cut-off: same option (-s slope), same
# stdout ("Significant hits:" / "Insignificant hits:") and the same `CUTOFF=` line on stderr as the reference
# (/root/reference/bin/ppp_cutoff.pl:91-162).  The rule: printed log-scores <= -0.001 are collected, sorted descending,
# scores above -1 ignored; the first adjacent pair whose relative difference exceeds the slope (in percent) fixes the cut-off at
# the earlier one; rows with a score below 0 are printed, the "Insignificant" marker before the first row above the cut-off.
# This step has no hot path: it is pure text filtering.  (For a threshold run that is still on the device the same cut-off is
# available without a text round trip: PhyloProf::B200::printed_cutoffs / ppp_printed_cutoffs.)
use strict;
use warnings;
no warnings 'numeric';
use Getopt::Long;

my ( $help, $slope );
GetOptions( 'help|?' => \$help, 'slope|s=i' => \$slope ) or die "usage: ppp_cutoff_b200.pl -s <slope percent> [table]\n";
die "usage: ppp_cutoff_b200.pl -s <slope percent> [table]\n" if $help || !defined $slope;
my $file = shift @ARGV;
my $fh;
if ( $file && -s $file ) { open( $fh, "<", $file ) or die "Can't open $file\n" }
else                     { $fh = \*STDIN }

my ( $header, @rows, @scores );
while ( my $line = <$fh> ) {
    unless ( defined $header ) { $header = $line; next }
    chomp $line;
    next if $line =~ /^-+$/;    # the -m marker line of ppp.pl
    push @rows, $line;
    next if $line =~ /^\#/;
    my $s = field5($line);
    push @scores, $s unless $s > -0.001;
}
close($fh);

my ( $last, $cutoff );
foreach my $s ( sort { $b <=> $a } @scores ) {
    next if $s > -1;
    unless ($last) { $last = $s; next }
    if ( abs( ( $last - $s ) / $last * 100 ) > $slope ) { $cutoff = $last; last }
    $last = $s;
}
print STDERR "CUTOFF=", ( defined $cutoff ? $cutoff : "" ), "\n";
print $header if defined $header;
print "Significant hits:\n";
my $marked = 0;
foreach my $line (@rows) {
    my $s = field5($line);
    next if $s >= 0;
    if ( $s > ( defined $cutoff ? $cutoff : 0 ) && !$marked ) { print "Insignificant hits:\n"; $marked = 1 }
    print $line, "\n";
}
exit(0);

sub field5 {    # column 6 as printed (compared numerically; CUTOFF= echoes the printed form); a missing field counts as 0
    my @f = split( /\t/, $_[0] );
    return defined $f[5] ? $f[5] : 0;
}
