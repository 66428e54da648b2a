#!/usr/bin/perl
# oracle/time_reference.pl -- TEST / BENCH INFRASTRUCTURE ONLY.
# Times the UNMODIFIED reference scorer (lib/PhyloProf/Algorithm/ppp.pm:130-245 under $REFERENCE_ROOT/lib, with
# Math::Cephes::bdtrc bound to the reference's own bundled binary through the XS shim) on a sample of a synthetic
# workload, one forked perl process per host core as bin/ppp.pl:895-922 forks its clients.  bench.py's cpu_baseline leg
# and `bench.py --impl reference` call it; the product never does.
#
#   perl time_reference.pl <cases.json> <nproc> <budget seconds> <out.json>
# cases.json: { "queries": [ {taxon: 0|1, ...}, ... ], "probs": [p, ...], "targets": [ [taxon, ...], ... ] }
# Every (query, target) pair is one get_match_score call; worker w takes pairs w, w + nproc, ...  Objects are built before
# the clock starts (the reference's drivers also build profiles outside the sweep); a worker stops early when its budget
# is spent.  out.json: { "pairs": n, "seconds": wall of the slowest worker, "workers": nproc,
#                        "results": { "q,t": [score %.17g, yes, number, depth] } }   (for the parity cross-check)
use strict;
use warnings;
use JSON::PP;
use FindBin;
use Time::HiRes qw(time);
BEGIN {
    my $root = $ENV{REFERENCE_ROOT} || "/root/reference";
    unshift @INC, "$FindBin::Bin/stubs", "$root/lib";
    $ENV{CEPHES_SHIM_SO} ||= "$FindBin::Bin/_ref/Cephes_shim.so";
    $ENV{CEPHES_REAL_SO} ||= "$FindBin::Bin/_ref/Cephes.so";
}
use PhyloProf::Profile;
use PhyloProf::Algorithm::ppp;

my ( $file, $nproc, $budget, $outfile ) = @ARGV;
die "usage: time_reference.pl cases.json nproc budget out.json\n" unless $outfile;
open( my $fh, "<", $file ) or die "$file: $!";
my $c = decode_json( do { local $/; <$fh> } );
close $fh;
my @q = map { PhyloProf::Profile->new( -profile => $_ ) } @{ $c->{queries} };
my @t = map { PhyloProf::Profile->new( -raw => $_, -yes => scalar(@$_), -no => 0 ) } @{ $c->{targets} };
my ( $nq, $nt ) = ( scalar @q, scalar @t );
my @pipes;
for my $w ( 0 .. $nproc - 1 ) {
    pipe( my $rd, my $wr ) or die;
    my $pid = fork();
    die "fork: $!" unless defined $pid;
    if ( !$pid ) {
        close $rd;
        open( STDOUT, ">", "/dev/null" );    # Cephes' mtherr prints on stdout
        my ( $n, %res ) = (0);
        my $t0 = time();
        for ( my $i = $w ; $i < $nq * $nt ; $i += $nproc ) {
            my ( $qi, $ti ) = ( int( $i / $nt ), $i % $nt );
            my @r = PhyloProf::Algorithm::ppp->new( -prof1 => $q[$qi], -prof2 => $t[$ti], -prob => $c->{probs}[$qi] )->get_match_score();
            $res{"$qi,$ti"} = [ sprintf( "%.17g", $r[0] ), $r[1] + 0, $r[2] + 0, $r[3] + 0 ];
            $n++;
            last if time() - $t0 > $budget;
        }
        my $dt = time() - $t0;
        print $wr encode_json( { pairs => $n, seconds => $dt, results => \%res } ), "\n";
        close $wr;
        exit(0);
    }
    close $wr;
    push @pipes, $rd;
}
my ( $pairs, $secs, %all ) = ( 0, 0 );
foreach my $rd (@pipes) {
    my $line = <$rd>;
    close $rd;
    next unless $line;
    my $r = decode_json($line);
    $pairs += $r->{pairs};
    $secs = $r->{seconds} if $r->{seconds} > $secs;
    @all{ keys %{ $r->{results} } } = values %{ $r->{results} };
}
1 while wait() != -1;
open( my $out, ">", $outfile ) or die "$outfile: $!";
print $out encode_json( { pairs => $pairs, seconds => $secs, workers => $nproc + 0, results => \%all } ), "\n";
close $out;
